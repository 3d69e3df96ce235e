#!/bin/bash
# One GPU-box pass: parity tests, tcgen05 descriptor probe, short benches.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests/test_lr_gpu.py tests/test_generator_gpu.py -q -m gpu > gpurun_out/pytest_fp32.log 2>&1
echo "fp32+lr tests rc=$?"; tail -5 gpurun_out/pytest_fp32.log
rm -f gpurun_out/probe_umma.json
for v in 0 1 2 3; do
  timeout 120 python tools/probe_umma.py --variant $v > gpurun_out/probe_v$v.log 2>&1
  echo "probe variant $v rc=$? ok=$(grep -c '"ok": true' gpurun_out/probe_v$v.log) bad=$(grep -c '"ok": false' gpurun_out/probe_v$v.log)"
done
timeout 900 python -m pytest tests/test_tc_gpu.py -q -m gpu > gpurun_out/pytest_tc.log 2>&1
echo "tc tests rc=$?"; tail -15 gpurun_out/pytest_tc.log
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 > gpurun_out/bench_fp32.log 2>&1
echo "bench fp32 rc=$?"; tail -1 gpurun_out/bench_fp32.log
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1
echo "bench bf16 rc=$?"; tail -2 gpurun_out/bench_bf16.log
