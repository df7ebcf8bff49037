#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "dwconv and fp16" > gpurun_out/pytest_k.log 2>&1; rc=$?; echo "dw tests rc=$rc"; tail -2 gpurun_out/pytest_k.log
for ring in 0 60 36; do
DN_DWS_RING_KB=$ring timeout 300 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_k_r$ring.json 2> gpurun_out/bench_k_r$ring.err; echo "bench ring$ring rc=$?"
done
python - <<'PY'
import json
for f in ("r0","r60","r36"):
    j=json.load(open("gpurun_out/bench_k_%s.json"%f))
    pk=j["roofline"]["per_kernel"]
    print(f, round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms; sync", round(j["api_list"]["engine_forward_synchronous"]["value"],1), "| dw", pk["dwconv kernels (stream / stream2 / tma / direct)"]["ms"])
PY
paste <(grep " dw " gpurun_out/bench_k_r0.err | cut -c1-52) <(grep " dw " gpurun_out/bench_k_r60.err | cut -c40-52) <(grep " dw " gpurun_out/bench_k_r36.err | cut -c40-52) | head -20
