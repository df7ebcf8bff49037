#!/bin/bash
# Runs on the GPU box under gpurun: GPU tests file by file (a hung kernel only costs its own timeout),
# then the bench.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1 t=$2; shift 2
  timeout $t python -m pytest "$@" -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? $(tail -1 gpurun_out/$name.log)"
}
run nms 900 tests/test_nms_gpu.py -m gpu
run post 900 tests/test_postprocess_gpu.py -m gpu
run kern_misc 600 tests/test_kernels_gpu.py -m gpu -k "not pwconv"
run kern_simt 600 tests/test_kernels_gpu.py -m gpu -k "pwconv_simt"
run kern_tc 600 tests/test_kernels_gpu.py -m gpu -k "pwconv_tc or pwconv_head"
run engine_simt 900 tests/test_engine_gpu.py -m gpu -k "simt"
run engine 1200 tests/test_engine_gpu.py -m gpu -k "not simt"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 10 --warmup 3 --layers > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit=$?"; tail -c 3000 gpurun_out/bench.log
