#!/bin/bash
# A/B of the GEMM tile-width cap (TMEM columns -> CTAs per SM) on one box
mkdir -p gpurun_out
for bn in 256 128 64; do
  DN_PW_BN_MAX=$bn timeout 300 python bench.py --steps 20 --warmup 5 --layers --no-cpu-baseline > gpurun_out/bench_f_bn$bn.json 2> gpurun_out/bench_f_bn$bn.err; echo "bn$bn rc=$?"
done
python - <<'PY'
import json
for f in ("bn256","bn128","bn64"):
    try:
        j=json.load(open("gpurun_out/bench_f_%s.json"%f))
        pk=j["roofline"]["per_kernel"]
        print(f, round(j["value"],1), j["unit"], round(j["ms_per_step"],3), "ms; sync", round(j["api_list"]["engine_forward_synchronous"]["value"],1), "| pw", pk["pwconv_tc_kernel"]["ms"], "sum", j["roofline"]["timing"][-60:])
    except Exception as e: print(f, "failed", e)
PY
paste <(grep " pw " gpurun_out/bench_f_bn256.err | cut -c1-52) <(grep " pw " gpurun_out/bench_f_bn128.err | cut -c40-52) <(grep " pw " gpurun_out/bench_f_bn64.err | cut -c40-52)
