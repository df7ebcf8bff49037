#!/bin/bash
# quick GPU round: post-processing + engine tests, bench, ncu launch list
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  timeout $t python -m pytest "$@" -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? $(tail -1 gpurun_out/$name.log)"; }
run nms 600 tests/test_nms_gpu.py -m gpu
run post 900 tests/test_postprocess_gpu.py -m gpu
run engine 1200 tests/test_engine_gpu.py -m gpu
timeout 900 python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit=$?"; tail -c 2500 gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "== ncu exit=$?"
