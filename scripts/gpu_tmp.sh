timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "pwconv or se_folded" 2>&1 | tail -2
DN_PW_PAIR=2 timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "benchmarked or layerwise or pipeline_mode" 2>&1 | tail -2
for p in 0 2 0 2; do
  DN_PW_PAIR=$p timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-extras | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pair=$p', round(d['value']), round(d['ms_per_step'],4))"
done
