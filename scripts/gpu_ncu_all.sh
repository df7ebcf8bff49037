#!/bin/bash
# One forward (graph replay, one batch in flight), EVERY kernel, with the ncu sections that say what bounds it.
# usage: gpu_ncu_all.sh <kernels per forward> [extra bench args]
N=${1:-110}
shift
mkdir -p gpurun_out
timeout 1500 ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats \
  --section MemoryWorkloadAnalysis --clock-control none --kernel-name-base demangled \
  -k 'regex:pwconv|pwdw_fused|dwpw_fused|dwconv|stem_|se_pool|se_fc|se_scale|softmax_decode|pick_thresholds|class_sort|class_nms|merge_topd' \
  -s $((2 * N)) -c $N --csv --page raw --log-file gpurun_out/ncu_all.csv python bench.py --pipeline 1 --no-extras --steps 1 --warmup 3 "$@" > gpurun_out/ncu_all_bench.log 2>&1
echo "== ncu exit=$?"
python scripts/ncu_all_summary.py gpurun_out/ncu_all.csv > gpurun_out/ncu_all_summary.txt; head -5 gpurun_out/ncu_all_summary.txt
