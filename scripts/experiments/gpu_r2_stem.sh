#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "test_stem" 2>&1 | tail -8 | tee gpurun_out/stem_pytest.log
for impl in simt tc; do
  DN_STEM=$impl timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/stem_$impl.json 2> gpurun_out/stem_$impl.err
  python - <<PY
import json
j=json.load(open("gpurun_out/stem_$impl.json"))
pk=j["roofline"]["per_kernel"]
print("$impl: %.0f img/s %.3f ms | "%(j["value"], j["ms_per_step"]) + " ".join("%s %.3f"%(k.split()[0][:14], v["ms"]) for k,v in pk.items()))
PY
done
