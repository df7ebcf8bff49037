#!/bin/bash
# End-of-iteration GPU pass: all GPU tests, smoke, bench, the other configs, and the ncu launch list with DRAM bytes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit=$?"
timeout 600 python scripts/bench_configs.py --config 4 > gpurun_out/config4.json 2> gpurun_out/config4.err; echo "== config4 exit=$? $(cut -c1-300 gpurun_out/config4.json)"
timeout 900 python scripts/bench_configs.py --config 5 > gpurun_out/config5.json 2> gpurun_out/config5.err; echo "== config5 exit=$? $(cut -c1-300 gpurun_out/config5.json)"
# one forward = 104 kernels of ours; skip the first two forwards (eager + capture), list the third
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:pwconv|dwconv|stem_conv|se_pool|se_fc|se_scale|softmax_decode|pick_thresholds|class_sort|class_nms|merge_topd' -s 208 -c 104 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "== ncu exit=$?"
