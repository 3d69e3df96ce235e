#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tc_gpu.py tests/test_generator_gpu.py -q -m gpu -x > gpurun_out/pytest_tc.log 2>&1
echo "tc+gen tests rc=$?"; tail -3 gpurun_out/pytest_tc.log
for pdl in 0 1; do
VTTS_PDL=$pdl timeout 600 python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.log 2>&1
echo "PDL=$pdl"; tail -1 gpurun_out/bench_pdl$pdl.log
done
