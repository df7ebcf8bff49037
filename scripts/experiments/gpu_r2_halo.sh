#!/bin/bash
mkdir -p gpurun_out
for m in 1 2; do
  echo "== DN_C3_HALO=$m"
  DN_C3_HALO=$m timeout 600 python -m pytest tests/test_vgg_kernels_gpu.py -m gpu -q -k "conv3x3" 2>&1 | tail -4
done
echo "== DN_C3_HALO=0"; DN_C3_HALO=0 timeout 600 python -m pytest tests/test_vgg_kernels_gpu.py -m gpu -q 2>&1 | tail -2
for m in 0 1; do
  DN_C3_HALO=$m timeout 600 python bench.py --config 6 --no-extras --steps 10 --warmup 3 > gpurun_out/bench_c6_halo$m.json 2> gpurun_out/bench_c6_halo$m.err; echo "bench halo=$m rc=$?"
  python - <<PY
import json
j=json.load(open("gpurun_out/bench_c6_halo$m.json"))
print(round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms/step roof", j["roofline"]["achieved"], j["roofline"]["frac"])
PY
done
