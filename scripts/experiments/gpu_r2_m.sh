#!/bin/bash
mkdir -p gpurun_out
for cb in 0 40 8; do
DN_DWS_CB=$cb timeout 300 python bench.py --steps 10 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_m_cb$cb.json 2> gpurun_out/bench_m_cb$cb.err; echo "bench cb$cb rc=$?"
done
paste <(grep " dw " gpurun_out/bench_m_cb0.err | cut -c1-52) <(grep " dw " gpurun_out/bench_m_cb40.err | cut -c40-52) <(grep " dw " gpurun_out/bench_m_cb8.err | cut -c40-52) | head -16
