#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config 6 --steps 10 --warmup 3 > gpurun_out/bench_c6.json 2> gpurun_out/bench_c6.err; echo "bench c6 rc=$?"; tail -3 gpurun_out/bench_c6.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_c6.json"))
print(round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms/step; e2e", round(j["e2e"]["value"],1), "roof", j["roofline"]["achieved"], j["roofline"]["frac"], "cpu", j.get("cpu_baseline",{}).get("value"), j["config"]["gflop_per_image"])
PY
# per-kernel view of one forward (B = 64)
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled \
  -k 'regex:conv3x3|pwconv|maxpool|l2norm|im2col' -s 38 -c 38 --csv --page raw --log-file gpurun_out/ncu_vgg.csv python bench.py --config 6 --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_vgg.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,re
rows=list(csv.reader([l for l in open("gpurun_out/ncu_vgg.csv") if not l.startswith("==")]))
h=rows[0]; ix={n:i for i,n in enumerate(h)}
tot=0
for r in rows[2:]:
    name=re.sub(r"\(.*","",r[ix["Kernel Name"]]).replace("void ","").replace("dn::","")[:40]
    us=float(r[ix["gpu__time_duration.sum"]].replace(",",""))/ (1e3 if rows[1][ix["gpu__time_duration.sum"]] in ("ns","nsecond") else 1)
    tot+=us
    print("%-40s %9.1f us  tensor %5s %%  grid %s" % (name, us, r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]], r[ix["Grid Size"]]))
print("total", tot)
PY
