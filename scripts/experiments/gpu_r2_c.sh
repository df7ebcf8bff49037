#!/bin/bash
# round 2: re-run the engine / C++ / transform tests, smoke, bench config 2 + ncu launch list + sanitizer on the postprocess and SE-pool paths
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu -s --deselect tests/test_nms_engine_path_gpu.py::test_config4_all_1024_images_bit_exact > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_c.log
grep -E "passed|failed" gpurun_out/pytest_c.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_c.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_c.log | cut -c1-600
timeout 400 python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench_c2_fp16.json 2> gpurun_out/bench_c2_fp16.err; echo "bench c2 fp16 rc=$?"
python - <<'PY'
import json
for f in ("c2_fp16",):
    try:
        j=json.load(open("gpurun_out/bench_%s.json"%f))
        print(f, round(j["value"],2), j["unit"], round(j["ms_per_step"],3), "ms; e2e", round(j.get("e2e",{}).get("value",0),1), "api", round(j.get("api_list",{}).get("value",0),1), "sync", round(j.get("api_list",{}).get("engine_forward_synchronous",{}).get("value",0),1), "roof", round(j.get("roofline",{}).get("frac",0),3))
    except Exception as e: print(f, "failed", e)
PY
# compute-sanitizer (memcheck + racecheck) on the kernels added / changed since the last sanitised version: POOL variants of the row streams, scored entry, postprocess
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "pooled_by_the_row_stream and fp16" > gpurun_out/sanitizer_memcheck_pool.log 2>&1; echo "memcheck pool rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_pool.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "pooled_by_the_row_stream and fp16 and (64-40 or 48-20 or 32-80 or 2-10)" > gpurun_out/sanitizer_racecheck_pool.log 2>&1; echo "racecheck pool rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_pool.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_nms_engine_path_gpu.py -q -m gpu -k "golden or deep_rounds or legacy or edge" > gpurun_out/sanitizer_memcheck_post.log 2>&1; echo "memcheck post rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_post.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_nms_engine_path_gpu.py -q -m gpu -k "golden_stress or deep_rounds" > gpurun_out/sanitizer_racecheck_post.log 2>&1; echo "racecheck post rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_post.log
