#!/bin/bash
# End-of-iteration GPU pass: all GPU tests, smoke, bench (+ reference arm), the other configs, the ncu launch list
# with DRAM bytes of one graph-replayed forward.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "== reference arm exit=$? $(cut -c1-300 gpurun_out/bench_ref.log)"
timeout 600 python scripts/bench_configs.py --config 4 > gpurun_out/config4.json 2> gpurun_out/config4.err; echo "== config4 exit=$? $(cut -c1-200 gpurun_out/config4.json)"
timeout 900 python scripts/bench_configs.py --config 5 > gpurun_out/config5.json 2> gpurun_out/config5.err; echo "== config5 exit=$? $(cut -c1-200 gpurun_out/config5.json)"
timeout 600 python scripts/layers_config5.py 512 > gpurun_out/layers_c5.txt 2>&1
bash scripts/gpu_launchlist.sh 118
