#!/bin/bash
# column strips of the depthwise row streams: parity tests, then config 5 (V2 @ 512) A/B on one box, then config 2 (must be unchanged)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/strips_pytest.log
for v in 1 0; do
  DN_DW_STRIPS=$v timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline --layers > gpurun_out/strips_c5_$v.json 2> gpurun_out/strips_c5_$v.err
  python - <<PY
import json
j=json.load(open("gpurun_out/strips_c5_$v.json"))
print("strips=$v", j["value"], j["ms_per_step"], {k:(v["ms"],v["frac_of_hbm_peak"]) for k,v in j["roofline"]["per_kernel"].items()})
PY
  grep " dw " gpurun_out/strips_c5_$v.err | head -4
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config2', d['value'], d['ms_per_step'])"
