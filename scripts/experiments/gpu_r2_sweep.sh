#!/bin/bash
# batch-size sweep of the pipelined engine: does a smaller batch (intermediates closer to the 126 MB L2) change img/s?
mkdir -p gpurun_out
for b in 32 64 128 256 512; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --batch $b 2> gpurun_out/sweep_b$b.err | tee gpurun_out/sweep_b$b.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B', $b, d['value'], d['ms_per_step'])"
done
