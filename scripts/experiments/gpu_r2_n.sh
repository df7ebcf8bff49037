#!/bin/bash
mkdir -p gpurun_out
for n in 148 111 74; do
DN_SM_COUNT=$n timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n_$n.json 2> gpurun_out/bench_n_$n.err; echo "bench sm$n rc=$?"
python -c "import json;j=json.load(open('gpurun_out/bench_n_$n.json'));print($n, round(j['value'],1), round(j['ms_per_step'],3))"
done
