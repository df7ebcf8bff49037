#!/bin/bash
# run given pytest targets (-m gpu), then the bench with the layer table
mkdir -p gpurun_out
timeout 1200 python -m pytest "$@" -m gpu -q --tb=short --maxfail=10 -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "== tests exit=$? $(tail -1 gpurun_out/t.log)"; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/t.log | head -30
timeout 600 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/ab_A.log 2> gpurun_out/ab_A.err; python -c "
import json
d=json.loads(open('gpurun_out/ab_A.log').readline()); print('value %.0f img/s %.3f ms  e2e %s' % (d['value'], d['ms_per_step'], {k:(round(v) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','ms_per_step')}), d.get('e2e_u8')); print({k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})"
