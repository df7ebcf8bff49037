#!/bin/bash
# full GPU test suite, then the bench under two environment settings.  usage: gpu_ab2.sh "<ENV for A>" "<ENV for B>"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider > gpurun_out/ab_tests.log 2>&1
echo "== tests exit=$? $(tail -1 gpurun_out/ab_tests.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/ab_tests.log | head
for arm in A B; do
  envs="$1"; [ $arm = B ] && envs="$2"
  env $envs timeout 600 python bench.py --steps 30 --warmup 5 --layers --no-cpu-baseline > gpurun_out/ab_$arm.log 2> gpurun_out/ab_$arm.err
  echo "== arm $arm ($envs) exit=$? $(python -c "
import json,sys
d=json.loads(open('gpurun_out/ab_$arm.log').readline())
print('value %.0f img/s  %.3f ms/step  e2e %.0f  e2e_u8 %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_u8']['value']))
print({k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})
")"
done
