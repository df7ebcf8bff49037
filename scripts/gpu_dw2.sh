#!/bin/bash
mkdir -p gpurun_out
DN_DW_IMPL=8 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k dwconv > gpurun_out/dw2_tests.log 2>&1; echo "== forced stream2 tests exit=$? $(tail -1 gpurun_out/dw2_tests.log)"; grep -E "^FAILED|max err" gpurun_out/dw2_tests.log | head
for arm in A B; do
  envs="DN_DW_IMPL=0"; [ $arm = B ] && envs="DN_DW_IMPL=8"
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/ab_$arm.log 2> gpurun_out/ab_$arm.err
  echo "== arm $arm ($envs) $(python -c "
import json
d=json.loads(open('gpurun_out/ab_$arm.log').readline()); print('value %.0f img/s %.3f ms' % (d['value'], d['ms_per_step']), d['roofline']['per_kernel']['dwconv_kernel']['ms'])")"
done
paste <(grep " dw " gpurun_out/ab_A.err | awk '{print $1,$2,$3,$4,$5,$6,$7}') <(grep " dw " gpurun_out/ab_B.err | awk '{print $7}') | grep " s2 "
