#!/bin/bash
# the other BASELINE configurations with the v5 build (four batches in flight by default): 4 (post-processing stress), 5 (V2 @ 512), 6 (VGG)
mkdir -p gpurun_out
for c in 4 5 6; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --layers > gpurun_out/v5_config$c.json 2> gpurun_out/v5_config$c.err; echo "config $c rc=$?"
  python - <<PY
import json
j=json.load(open("gpurun_out/v5_config$c.json"))
print("config $c", round(j["value"],3), j["unit"], round(j["ms_per_step"],3), "e2e", j.get("e2e",{}).get("value"), j["config"].get("l2"), {k:(v["ms"],v["frac_of_hbm_peak"]) for k,v in j.get("roofline",{}).get("per_kernel",{}).items()})
PY
done
