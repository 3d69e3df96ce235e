#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -q -m gpu > gpurun_out/pytest_tc.log 2>&1
echo "tc tests rc=$?"; tail -12 gpurun_out/pytest_tc.log
VTTS_PROFILE=1 timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_prof.log 2>&1
echo "bench prof rc=$?"; grep "vtts-prof" gpurun_out/bench_bf16_prof.log | tail -82 | awk '{print}' | head -90
timeout 600 python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1
echo "bench bf16 rc=$?"; tail -1 gpurun_out/bench_bf16.log
