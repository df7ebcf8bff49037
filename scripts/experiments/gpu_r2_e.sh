#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "se_folded" > gpurun_out/pytest_sefold.log 2>&1; rc=$?; echo "se-fold kernel test rc=$rc"; tail -2 gpurun_out/pytest_sefold.log
if [ $rc -ne 0 ]; then grep -E "^E  " gpurun_out/pytest_sefold.log | head -10; fi
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s > gpurun_out/pytest_e.log 2>&1; echo "engine tests rc=$?"; grep -E "passed|failed" gpurun_out/pytest_e.log | tail -2; grep -E "^FAILED" gpurun_out/pytest_e.log | head
timeout 400 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_e_fold.json 2> gpurun_out/bench_e_fold.err; echo "bench fold rc=$?"
DN_SE_FOLD=0 timeout 400 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_e_nofold.json 2> gpurun_out/bench_e_nofold.err; echo "bench nofold rc=$?"
python - <<'PY'
import json
for f in ("e_fold","e_nofold"):
    try:
        j=json.load(open("gpurun_out/bench_%s.json"%f))
        pk=j["roofline"]["per_kernel"]
        print(f, round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms; api", round(j["api_list"]["value"],1), "sync", round(j["api_list"]["engine_forward_synchronous"]["value"],1), "| pw", pk["pwconv_tc_kernel"]["ms"], "se", pk["se kernels (fc1 + fc2 + scale)"]["ms"], "sum", j["roofline"]["timing"][-60:])
    except Exception as e: print(f, "failed", e)
PY
grep -E "^ *(12|16|20|36|40|44|48|52) " gpurun_out/bench_e_fold.err
