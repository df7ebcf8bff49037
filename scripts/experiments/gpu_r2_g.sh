#!/bin/bash
export NCU_CMD="python bench.py --pipeline 1 --no-extras --steps 1 --warmup 3"
bash scripts/gpu_ncu.sh "pwconv_tc:109:pw_20_80_184" "pwconv_tc:110:pw_20_184_80_res" "dwconv_stream_kernel:29:dws_40_120_k5"
for t in pw_20_80_184 pw_20_184_80_res dws_40_120_k5; do python scripts/ncu_source_summary.py gpurun_out/ncu/$t.source.csv 40 > gpurun_out/ncu/$t.source_summary.txt 2>&1; rm -f gpurun_out/ncu/$t.source.csv; done
