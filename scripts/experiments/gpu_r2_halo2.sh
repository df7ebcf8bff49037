#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vgg_kernels_gpu.py tests/test_vgg_gpu.py -m gpu -q 2>&1 | tail -5
DN_C3_HALO=0 timeout 600 python bench.py --config 6 --no-extras --steps 10 --warmup 3 > gpurun_out/bench_c6_halo0.json 2> gpurun_out/bench_c6_halo0.err
python -c "
import json
j=json.load(open('gpurun_out/bench_c6_halo0.json')); print('halo off', round(j['value'],1), round(j['ms_per_step'],3))"
bash scripts/gpu_r2_vgg.sh
