# three 20-step bench runs (sustained clocks under the power cap): ms per step, SM clock, cycles per step
for rep in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 | python3 -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('run', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['ms_per_step']*d['clocks']['sm_mhz']/1000,2),'Mcycles', d['clocks']['reasons'])"
done
