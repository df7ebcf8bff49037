#!/bin/bash
# up to four batches in flight: engine tests, then config 2 with 2 / 3 / 4 slots on one box (with the e2e legs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_cpp_caller.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/slots_pytest.log
for n in 2 4 3 4 2; do
  timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --pipeline $n > gpurun_out/slots_$n.json 2> gpurun_out/slots_$n.err
  python - <<PY
import json
j=json.load(open("gpurun_out/slots_$n.json"))
print("slots=$n", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "u8", round(j["e2e_u8"]["value"]), "api", j.get("api_list",{}).get("value"), j["config"]["l2"])
PY
done
