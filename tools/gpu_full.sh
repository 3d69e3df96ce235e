#!/bin/bash
# Full GPU pass: all gpu tests, smoke, default bench, reference arm, ncu launch list (time + DRAM bytes) + full captures.
# One forward of tools/ncu_forward.py launches 52 kernels matching the regex below:
#   cf_to_cl, conv_tc x23 (pre, up0, 18 x stage 0, up1, up2, up3), unit_tc x9 (stage 1), unit64 x18 (stages 2, 3), conv_post
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
echo "bench reference rc=$?"; tail -1 gpurun_out/bench_reference.log
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
K='regex:conv_tc|unit_tc|unit64_tc|conv_post|cf_to_cl'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 52 -c 52 --csv --log-file gpurun_out/launches.csv python tools/ncu_forward.py > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:unit_tc_kernel -s 15 -c 1 -o gpurun_out/prof_unit_c128_k11 -f python tools/ncu_forward.py > gpurun_out/ncu_a.log 2>&1
echo "unit c128 k11 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:unit64_tc_kernel -s 24 -c 1 -o gpurun_out/prof_unit64_c64_k11 -f python tools/ncu_forward.py > gpurun_out/ncu_b.log 2>&1
echo "unit64 c64 k11 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:unit64_tc_kernel -s 27 -c 1 -o gpurun_out/prof_unit64_c32_k3 -f python tools/ncu_forward.py > gpurun_out/ncu_c.log 2>&1
echo "unit64 c32 k3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 37 -c 2 -o gpurun_out/prof_conv_c256_k11 -f python tools/ncu_forward.py > gpurun_out/ncu_d.log 2>&1
echo "conv c256 k11 rc=$?"
ls -la gpurun_out/*.ncu-rep
