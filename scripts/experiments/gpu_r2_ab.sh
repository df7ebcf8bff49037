#!/bin/bash
# A/B of variant library builds on ONE box: usage gpu_r2_ab.sh <libdir> <libdir> ...   ("lib" = the default build)
mkdir -p gpurun_out
for rep in 1 2; do
for d in "$@"; do
  export DN_LIB_DIR=$PWD/demonet_b200/$d
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --layers > gpurun_out/ab_$d.json 2> gpurun_out/ab_$d.err
  python - <<PY
import json
j=json.load(open("gpurun_out/ab_$d.json"))
pk=j["roofline"]["per_kernel"]
print("$d rep$rep: %.0f img/s %.3f ms | "%(j["value"], j["ms_per_step"]) + " ".join("%s %.3f"%(k.split()[0][:14], v["ms"]) for k,v in pk.items()))
PY
done
done
