#!/bin/bash
# round 2, first GPU pass: the whole -m gpu suite, smoke(), bench in both storage types
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_a.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_a.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench fp16 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --act-dtype bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?"
python - <<'PY'
import json
for dt in ("fp16","bf16"):
    try:
        j=json.load(open("gpurun_out/bench_%s.json"%dt)); print(dt, round(j["value"]), "img/s", round(j["ms_per_step"],3), "ms; e2e", round(j["e2e"]["value"]), "u8", round(j["e2e_u8"]["value"]))
    except Exception as e: print(dt, "failed", e)
PY
