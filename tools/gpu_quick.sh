#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -q -m gpu -x > gpurun_out/pytest_tc.log 2>&1
echo "tc tests rc=$?"; tail -3 gpurun_out/pytest_tc.log
VTTS_PROFILE=1 timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prof.log 2>&1
echo "bench prof rc=$?"; grep "vtts-prof" gpurun_out/bench_prof.log | tail -52 | sed "s/^/cl1 /"
timeout 600 python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.log 2>&1
tail -1 gpurun_out/bench_fp16.log
timeout 300 python tools/trace_unit.py ${TRACE_LAUNCHES:-42 48 32 38 28 30} > gpurun_out/trace_unit.log 2>&1
grep -E "^---|^period" gpurun_out/trace_unit.log
