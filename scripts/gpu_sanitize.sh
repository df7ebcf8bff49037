#!/bin/bash
# compute-sanitizer passes over small workloads: memcheck on the smoke forward, racecheck on the NMS / post-processing
# kernels (shared-memory hazards; the batched-NMS race of r01 was of that kind).
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck.log 2>&1; echo "== memcheck exit=$? $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/san_memcheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_nms_gpu.py tests/test_postprocess_gpu.py -m gpu -q -x -p no:cacheprovider -k "golden_cases or vs_oracle or idempotence or padded" > gpurun_out/san_racecheck.log 2>&1; echo "== racecheck exit=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/san_racecheck.log | tr '\n' ' ')"
grep -E "Race reported|hazard" gpurun_out/san_racecheck.log | sort | uniq -c | head -20
