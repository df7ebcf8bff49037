#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "pwconv or se_folded or fused" > gpurun_out/pytest_h.log 2>&1; echo "kernel tests rc=$?"; tail -2 gpurun_out/pytest_h.log
timeout 300 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_h.json"))
pk=j["roofline"]["per_kernel"]
print(round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms; sync", round(j["api_list"]["engine_forward_synchronous"]["value"],1), "api", round(j["api_list"]["value"],1), "| pw", pk["pwconv_tc_kernel"]["ms"], "dw", pk["dwconv kernels (stream / stream2 / tma / direct)"]["ms"], "sum", j["roofline"]["timing"][-60:])
PY
grep " pw " gpurun_out/bench_h.err | cut -c1-52 | head -48
