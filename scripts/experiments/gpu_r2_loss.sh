#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_cpp_caller.py tests/test_host_cpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/loss_pytest.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_loss_gpu.py -q -x -k "golden or edges or random_cases_vs_oracle and not 4-64" 2>&1 | tail -8 | tee gpurun_out/loss_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_loss_gpu.py -q -x -k "golden or edges" 2>&1 | tail -8 | tee gpurun_out/loss_racecheck.log
