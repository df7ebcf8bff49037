#!/bin/bash
# multi-GPU check of the four-slot pipeline: weak scaling (config 2) with e2e legs, strong scaling (config 3)
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/scale5_${tag}_n$N.json 2> gpurun_out/scale5_${tag}_n$N.err; echo "$tag rc=$?"; }
run full --steps 20 --warmup 5 --no-cpu-baseline
run g4 --steps 40 --warmup 8 --no-extras --gather-every 4
run c3 --config 3 --steps 20 --warmup 5 --no-extras
python - <<PY
import json
for t in ("full","g4","c3"):
    try:
        j=json.load(open("gpurun_out/scale5_%s_n$N.json"%t))
        print(t, round(j["value"],1), j["unit"], "ms/step", round(j["ms_per_step"],3), "e2e", round(j.get("e2e",{}).get("value",0),1), "u8", round(j.get("e2e",{}).get("u8",{}).get("value",0),1), j["config"]["batch_per_gpu"])
    except Exception as e: print(t, "failed", e)
PY
tail -3 gpurun_out/scale5_full_n$N.err
