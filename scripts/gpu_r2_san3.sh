#!/bin/bash
# compute-sanitizer over what changed in r02 v5: column strips of the depthwise row streams, the cluster form of the SE
# fully-connected layers, the four-slot pipeline (memcheck on a pipelined forward_batches run); then the ncu launch list of one
# graph-replayed forward of the v5 build
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "column_strips or cluster_form or (test_dwconv and (256 or 150 or 128 or 131 or 70))" > gpurun_out/san3_memcheck_strips.log 2>&1; echo "== memcheck strips/cluster exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san3_memcheck_strips.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "cluster_form or (column_strips and fp16)" > gpurun_out/san3_racecheck.log 2>&1; echo "== racecheck exit=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/san3_racecheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -p no:cacheprovider -k "pipeline_mode and 4" > gpurun_out/san3_memcheck_slots.log 2>&1; echo "== memcheck slots exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san3_memcheck_slots.log | tr '\n' ' ')"
bash scripts/gpu_launchlist.sh 110
