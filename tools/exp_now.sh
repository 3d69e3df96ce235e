mkdir -p gpurun_out
for v in 0 1; do
  echo "=== NO_W_RELOAD=$v"
  VTTS_DBG_NO_W_RELOAD=$v VTTS_PROFILE=1 timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "vtts-prof" | tail -52 | grep "kind=2\|total"
  VTTS_DBG_NO_W_RELOAD=$v timeout 300 python tools/trace_unit.py 42 48 32 38 28 2>&1 | grep -E "^---|^period"
done
