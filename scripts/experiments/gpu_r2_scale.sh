#!/bin/bash
# multi-GPU: weak scaling (config 2) with the grouped gather vs a collective per step, strong scaling (config 3), e2e legs
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
(numactl -H || lscpu | grep -i numa) > gpurun_out/numa_n$N.txt 2>&1
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err; echo "$tag rc=$?"; }
run g4 --steps 40 --warmup 8 --no-extras --gather-every 4
run g1 --steps 40 --warmup 8 --no-extras --gather-every 1
run full --steps 20 --warmup 5 --no-cpu-baseline
run c3 --config 3 --steps 20 --warmup 5 --no-extras
python - <<PY
import json
for t in ("g4","g1","full","c3"):
    try:
        j=json.load(open("gpurun_out/scale_%s_n$N.json"%t))
        print(t, round(j["value"],1), j["unit"], "ms/step", round(j["ms_per_step"],3), "e2e", round(j.get("e2e",{}).get("value",0),1), "u8", round(j.get("e2e",{}).get("u8",{}).get("value",0),1), j["config"].get("host_affinity"), j["config"]["batch_per_gpu"])
    except Exception as e: print(t, "failed", e)
PY
