#!/bin/bash
# full ncu captures (stall reasons + source) of the three worst bandwidth-bound kernels
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --pipeline 1" bash scripts/gpu_ncu.sh "dwconv_stream_kernel:29:dws_k5_40" "dwconv_stream_kernel:28:dws_k3_80"
for t in dws_k5_40 dws_k3_80; do python scripts/ncu_source_summary.py gpurun_out/ncu/$t.source.csv 12 > gpurun_out/ncu/$t.source_summary.txt; done
python - <<'PY'
import csv
for t in ("dws_k5_40","dws_k3_80"):
    print("==", t)
    for r in csv.reader(open("gpurun_out/ncu/%s.raw.csv"%t)):
        if len(r)<3: continue
        h,v=r[0],r[-1]
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try: x=float(v.replace(",",""))
            except: continue
            if x>150: print("   %-60s %s"%(h.replace("smsp__pcsamp_warps_issue_stalled_",""),v))
        elif any(w in h for w in ("gpu__time_duration.sum","issue_active.avg.pct","warps_active.avg.pct","smsp__inst_executed.sum","launch__grid_size","registers_per_thread ")):
            print("   %-60s %s"%(h,v))
    print(open("gpurun_out/ncu/%s.source_summary.txt"%t).read()[:1800])
PY
