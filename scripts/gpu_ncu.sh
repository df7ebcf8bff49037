#!/bin/bash
# ncu --set full captures of representative launches; only small CSV summaries come back.
# usage: gpu_ncu.sh "<kernel-regex>:<skip>:<tag>" ...
mkdir -p gpurun_out/ncu
CMD=${NCU_CMD:-"python bench.py --steps 1 --warmup 3 --no-cpu-baseline"}
KEEP='issue_stalled|pcsamp|dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__occupancy_limit|sm__pipe_tensor|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|smsp__warp_issue_stalled.*_per_warp_active.pct|smsp__inst_executed.sum|launch__grid_size|launch__block_size|launch__shared_mem|l1tex__data_bank_conflicts|smsp__cycles_active.avg|sm__inst_executed_pipe|lts__t_bytes.sum|l1tex__t_bytes.sum|smsp__issue_active.avg.pct|achieved_occupancy|sm__maximum_warps'
for spec in "$@"; do
  IFS=: read pat skip tag <<< "$spec"
  rm -f /tmp/p.ncu-rep
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o /tmp/p $CMD > gpurun_out/ncu/$tag.log 2>&1
  echo "== $tag exit=$?"
  ncu -i /tmp/p.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import sys,csv,re
rows=list(csv.reader(sys.stdin))
if len(rows)>=3:
    hdr,units,vals=rows[0],rows[1],rows[2]
    pat=re.compile(r'$KEEP')
    for h,u,v in zip(hdr,units,vals):
        if pat.search(h) or h in ('Kernel Name','Block Size','Grid Size'): print('%s,%s,%s'%(h,u,v))
" > gpurun_out/ncu/$tag.raw.csv
  ncu -i /tmp/p.ncu-rep --page source --csv 2>/dev/null > gpurun_out/ncu/$tag.source.csv
  ls -la /tmp/p.ncu-rep | awk '{print $5}'
done
du -sh gpurun_out
