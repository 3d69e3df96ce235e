# step-level A/B of the L2 evict-first policy on activation / residual reads (VTTS_TC_STREAM_HINT)
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_generator_gpu.py -q -m gpu -x 2>&1 | tail -1
for rep in 1 2 3; do
  for v in 0 1; do
    VTTS_TC_STREAM_HINT=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 | python3 -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('hint $v', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['ms_per_step']*d['clocks']['sm_mhz']/1000,2),'Mcycles', round(d['config']['generator_ms_untrimmed'],3))"
  done
done
