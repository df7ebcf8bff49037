#!/bin/bash
# round 2: the whole -m gpu suite (no -x), smoke(), bench for configs 2 (fp16 + bf16), 4, 5
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_b.log
grep -E "passed|failed" gpurun_out/pytest_b.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_b.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_b.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_b.log
timeout 400 python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench_c2_fp16.json 2> gpurun_out/bench_c2_fp16.err; echo "bench c2 fp16 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --act-dtype bf16 > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "bench c2 bf16 rc=$?"
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 400 python bench.py --config 5 --steps 10 --warmup 3 --layers > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
python - <<'PY'
import json
for f in ("c2_fp16","c2_bf16","c4","c5"):
    try:
        j=json.load(open("gpurun_out/bench_%s.json"%f))
        print(f, round(j["value"],2), j["unit"], round(j["ms_per_step"],3), "ms; e2e", round(j.get("e2e",{}).get("value",0),1), "api", round(j.get("api_list",{}).get("value",0),1), "roof", round(j.get("roofline",{}).get("frac",0),3))
    except Exception as e: print(f, "failed", e)
PY
