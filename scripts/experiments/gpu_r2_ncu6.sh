#!/bin/bash
# full ncu capture of the 10x10x480 5x5 depthwise launch (row stream, squeeze-excitation pooling on the side)
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --pipeline 1" bash scripts/gpu_ncu.sh "dwconv_stream_kernel:8:dws_k5_10"
python scripts/ncu_source_summary.py gpurun_out/ncu/dws_k5_10.source.csv 40 > gpurun_out/ncu/dws_k5_10.source_summary.txt
rm -f gpurun_out/ncu/dws_k5_10.source.csv
