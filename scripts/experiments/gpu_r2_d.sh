#!/bin/bash
# round 2: SE fold first (short timeout: a protocol bug would hang), then everything of gpu_r2_c.sh
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "se_folded" > gpurun_out/pytest_sefold.log 2>&1; rc=$?; echo "se-fold kernel test rc=$rc"; tail -4 gpurun_out/pytest_sefold.log
if [ $rc -ne 0 ]; then grep -E "^E  " gpurun_out/pytest_sefold.log | head -10; export DN_SE_FOLD=0; echo "SE fold DISABLED for the rest of this run"; fi
bash scripts/gpu_r2_c.sh
