#!/bin/bash
# full ncu captures of three GEMM launches: cls head level 0 (672 -> 546, fp32 out), SE-scaled project 40x40 120 -> 40, 20x20 672 -> 112
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --pipeline 1" bash scripts/gpu_ncu.sh "pwconv_tc_kernel:36:pw_cls_l0" "pwconv_tc_kernel:6:pw_se_40_120_40" "pwconv_tc_kernel:20:pw_se_20_672_112"
for t in pw_cls_l0 pw_se_40_120_40 pw_se_20_672_112; do python scripts/ncu_source_summary.py gpurun_out/ncu/$t.source.csv 14 > gpurun_out/ncu/$t.source_summary.txt; rm -f gpurun_out/ncu/$t.source.csv; done
head -5 gpurun_out/ncu/pw_cls_l0.raw.csv
