#!/bin/bash
# A/B on the GPU box: engine tests, then the bench with and without an environment toggle.
# usage: gpu_ab.sh "<ENV=VAL for arm A>" "<ENV=VAL for arm B>" [pytest -k expression]
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider ${3:+-k "$3"} > gpurun_out/ab_tests.log 2>&1
echo "== tests exit=$? $(tail -1 gpurun_out/ab_tests.log)"
for arm in A B; do
  envs="$1"; [ $arm = B ] && envs="$2"
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/ab_$arm.log 2> gpurun_out/ab_$arm.err
  echo "== arm $arm ($envs) exit=$? $(python -c "
import json,sys
d=json.loads(open('gpurun_out/ab_$arm.log').readline())
print('value %.0f img/s  %.3f ms/step  e2e %.0f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))
print({k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})
")"
done
