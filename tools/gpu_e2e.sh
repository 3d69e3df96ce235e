#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "synth or trim" > gpurun_out/pytest_tc.log 2>&1
echo "synth tests rc=$?"; tail -3 gpurun_out/pytest_tc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_e2e.log
