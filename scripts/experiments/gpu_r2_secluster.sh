#!/bin/bash
# SE fc1 + fc2 as one cluster launch: parity, then A/B of config 2 on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "se or engine or layerwise or benchmarked" 2>&1 | tail -5 | tee gpurun_out/secl_pytest.log
for v in 1 0 1 0; do
  DN_SE_CLUSTER=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cluster=$v', d['value'], d['ms_per_step'], d.get('gpu_launches_per_step'))"
done
