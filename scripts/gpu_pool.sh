#!/bin/bash
# SE pooling by the depthwise row stream: parity tests, memcheck of the new test, bench with and without it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider > gpurun_out/pool_tests.log 2>&1
echo "== tests exit=$? $(tail -1 gpurun_out/pool_tests.log)"
grep -E "FAILED|Error|assert" gpurun_out/pool_tests.log | head -20
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "dwconv_se and (37 or 96 or 16)" > gpurun_out/pool_memcheck.log 2>&1
echo "== memcheck exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/pool_memcheck.log | tr '\n' ' ')"
for arm in A B A2 B2; do
  envs="DN_SE_POOL=0"; [ ${arm:0:1} = B ] && envs="DN_SE_POOL=1"
  lay=""; [ $arm = B ] && lay="--layers"
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 $lay --no-cpu-baseline > gpurun_out/pool_$arm.log 2> gpurun_out/pool_$arm.err
  echo "== arm $arm ($envs) exit=$? $(python -c "
import json,sys
d=json.loads(open('gpurun_out/pool_$arm.log').readline())
print('value %.0f img/s  %.3f ms/step  e2e %.0f  launches/step %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches_per_step')))
")"
done
grep -E "^ *[0-9]+ (se|dw) " gpurun_out/pool_B.err | head -30
