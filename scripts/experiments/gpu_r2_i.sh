#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "(pwconv or se_folded or fused) and fp16" > gpurun_out/pytest_i.log 2>&1; rc=$?; echo "kernel tests rc=$rc"; tail -2 gpurun_out/pytest_i.log
if [ $rc -ne 0 ]; then grep -E "^E  " gpurun_out/pytest_i.log | head; exit 1; fi
for ws in 1 0; do
DN_PW_WSTAT=$ws timeout 300 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_i_ws$ws.json 2> gpurun_out/bench_i_ws$ws.err; echo "bench ws$ws rc=$?"
done
python - <<'PY'
import json
for f in ("ws1","ws0"):
    j=json.load(open("gpurun_out/bench_i_%s.json"%f))
    pk=j["roofline"]["per_kernel"]
    print(f, round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms; sync", round(j["api_list"]["engine_forward_synchronous"]["value"],1), "| pw", pk["pwconv_tc_kernel"]["ms"], "sum", j["roofline"]["timing"][-60:])
PY
paste <(grep " pw " gpurun_out/bench_i_ws1.err | cut -c1-52) <(grep " pw " gpurun_out/bench_i_ws0.err | cut -c40-52) | head -46
