# A/B of the operand/output epilogue warp split of the fused unit kernels (VTTS_UNIT_EA_WARPS / VTTS_UNIT64_EA_WARPS = 8 or 4)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "v1 or trim or stages or odd" 2>&1 | tail -1
for v in 8 4; do
  echo "=== VTTS_UNIT64_EA_WARPS=$v"
  VTTS_UNIT64_EA_WARPS=$v VTTS_PROFILE=1 timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | grep "vtts-prof" | grep "kind=2" | grep "L=  97152\|L= 194304" | tail -18
done
