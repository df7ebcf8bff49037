#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "dwconv and fp16" > gpurun_out/pytest_l.log 2>&1; rc=$?; echo "dw tests rc=$rc"; tail -2 gpurun_out/pytest_l.log
timeout 300 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_l.json"))
pk=j["roofline"]["per_kernel"]
print(round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms; sync", round(j["api_list"]["engine_forward_synchronous"]["value"],1), "api", round(j["api_list"]["value"],1), "| dw", pk["dwconv kernels (stream / stream2 / tma / direct)"]["ms"], "pw", pk["pwconv_tc_kernel"]["ms"])
PY
grep " dw " gpurun_out/bench_l.err | cut -c1-52 | head -16
