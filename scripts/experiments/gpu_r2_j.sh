#!/bin/bash
# full verification: all gpu tests (both builds), smoke, bench config 2 (with cpu baseline), configs 4, 5, bf16, launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/pytest_j.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest_j.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_j.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_j.log | cut -c1-400
timeout 400 python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench_j_c2.json 2> gpurun_out/bench_j_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --act-dtype bf16 > gpurun_out/bench_j_c2_bf16.json 2> gpurun_out/bench_j_c2_bf16.err; echo "bench c2 bf16 rc=$?"
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/bench_j_c4.json 2> gpurun_out/bench_j_c4.err; echo "bench c4 rc=$?"
timeout 400 python bench.py --config 5 --steps 10 --warmup 3 --layers > gpurun_out/bench_j_c5.json 2> gpurun_out/bench_j_c5.err; echo "bench c5 rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_j_ref.json 2> gpurun_out/bench_j_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
for f in ("c2","c2_bf16","c4","c5","ref"):
    try:
        j=json.load(open("gpurun_out/bench_j_%s.json"%f))
        print(f, round(j["value"],2), j["unit"], "ms/step", round(j["ms_per_step"],3), "e2e", round(j.get("e2e",{}).get("value",0),1), "u8", round(j.get("e2e",{}).get("u8",{}).get("value",0),1), "api", round(j.get("api_list",{}).get("value",0),1), "sync", round(j.get("api_list",{}).get("engine_forward_synchronous",{}).get("value",0),1), "roof", round(j.get("roofline",{}).get("frac",0),3), "cpu", j.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
N=$(python -c "import json;print(json.load(open('gpurun_out/bench_j_c2.json'))['gpu_launches_per_step'])")
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:pwconv|pwdw_fused|dwpw_fused|dwconv|stem_|se_pool|se_fc|se_scale|softmax_decode|pick_thresholds|class_sort|class_nms|merge_topd' \
  -s $((2 * N)) -c $N --csv --log-file gpurun_out/launches.csv python bench.py --pipeline 1 --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "== ncu exit=$? (N=$N)"
python scripts/ncu_launch_summary.py gpurun_out/launches.csv gpurun_out/launches.json | tail -22
