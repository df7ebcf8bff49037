#!/bin/bash
# ncu launch list (time + DRAM bytes) of ONE graph-replayed forward: skip the eager and the first replayed forward.
# usage: gpu_launchlist.sh <kernels per forward>
N=${1:-128}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:pwconv|pwdw_fused|dwpw_fused|dwconv|stem_|se_pool|se_fc|se_scale|softmax_decode|pick_thresholds|class_sort|class_nms|merge_topd' \
  -s $((2 * N)) -c $N --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu exit=$?"
python scripts/ncu_launch_summary.py gpurun_out/launches.csv gpurun_out/launches.json
