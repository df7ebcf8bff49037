#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== PDL on"; timeout 300 python scripts/exp_pool.py 2>&1 | tail -30
timeout 600 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "dwconv_se and (37 or 96 or 16)" > gpurun_out/pool_memcheck.log 2>&1
echo "== memcheck exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/pool_memcheck.log | tr '\n' ' ')"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "dwconv_se and (37 or 16-10)" > gpurun_out/pool_racecheck.log 2>&1
echo "== racecheck exit=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/pool_racecheck.log | tr '\n' ' ')"
