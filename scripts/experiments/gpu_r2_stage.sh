#!/bin/bash
# tap staging of the row streams moved in front of the PDL wait, batched loads: parity tests, per-layer table, step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dwconv" 2>&1 | tail -3 | tee gpurun_out/stage_pytest.log
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --layers > gpurun_out/stage_bench.json 2> gpurun_out/stage_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/stage_bench.json"))
print(round(j["value"]), round(j["ms_per_step"],3), {k:v["ms"] for k,v in j["roofline"]["per_kernel"].items()})
PY
grep " dw " gpurun_out/stage_bench.err | head -40
