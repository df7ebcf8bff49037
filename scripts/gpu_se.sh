#!/bin/bash
# SE check: tests, per-kernel ncu times of the SE launches, bench with the layer table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "se or layerwise or end_to_end" > gpurun_out/se_tests.log 2>&1; echo "== tests exit=$? $(tail -1 gpurun_out/se_tests.log)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:se_pool|se_fc|se_scale' -c 32 --csv --log-file gpurun_out/se_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/se_ncu.log 2>&1
grep -E "se_(pool|fc|scale)" gpurun_out/se_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-14,120- | tr -d '"' | paste - - - | head -8
timeout 600 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/ab_A.log 2> gpurun_out/ab_A.err; python -c "
import json
d=json.loads(open('gpurun_out/ab_A.log').readline()); print('value %.0f img/s %.3f ms' % (d['value'], d['ms_per_step'])); print({k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})"
grep " se " gpurun_out/ab_A.err
