#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vgg_kernels_gpu.py tests/test_vgg_gpu.py -m gpu -x -q 2>&1 | tail -5
bash scripts/gpu_r2_vgg.sh
