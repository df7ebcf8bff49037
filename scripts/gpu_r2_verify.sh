#!/bin/bash
# full verification of HEAD: GPU tests, smoke, default bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/verify_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/verify_smoke.log
timeout 600 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/verify_bench.json"))
print(j["value"], j["ms_per_step"], j["e2e"], j["roofline"], j.get("clocks"))
PY
