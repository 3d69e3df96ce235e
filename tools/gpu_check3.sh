#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -q -m gpu > gpurun_out/pytest_tc.log 2>&1
echo "tc tests rc=$?"; tail -3 gpurun_out/pytest_tc.log
for cl in 1 0; do
  VTTS_TC_CLUSTER=$cl VTTS_PROFILE=1 timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prof_cl$cl.log 2>&1
  echo "bench prof cluster=$cl rc=$?"; grep "vtts-prof" gpurun_out/bench_prof_cl$cl.log | tail -78 | sed "s/^/cl$cl /"
  tail -1 gpurun_out/bench_prof_cl$cl.log | cut -c1-200
done
