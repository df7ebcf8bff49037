#!/bin/bash
# compute-sanitizer over the kernels changed in this session: tensor-core stem, dwpw_fused at three CTAs per SM (smoke forward =
# the whole V3 plan), the stem kernel tests (ring wrap), training-side kernels
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san2_memcheck_smoke.log 2>&1; echo "== memcheck smoke exit=$? $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/san2_memcheck_smoke.log | cut -c1-80 | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "test_stem or fused" > gpurun_out/san2_memcheck_stem.log 2>&1; echo "== memcheck stem/fused exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san2_memcheck_stem.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "test_stem and tc and 40" > gpurun_out/san2_racecheck_stem.log 2>&1; echo "== racecheck stem exit=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/san2_racecheck_stem.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "test_stem and tc and 40" > gpurun_out/san2_synccheck_stem.log 2>&1; echo "== synccheck stem exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san2_synccheck_stem.log | tr '\n' ' ')"
