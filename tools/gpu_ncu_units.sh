#!/bin/bash
# ncu full captures: wide unit (stage 1, k=11 d=1: 7th unit_tc launch of the 2nd forward) and narrow units of stage 2/3
mkdir -p gpurun_out
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -1 gpurun_out/ncu_plain.log
ncu --set full --clock-control none --import-source on -k regex:unit_tc_kernel -s 15 -c 1 -o gpurun_out/prof_u128_k11 -f python tools/ncu_forward.py > gpurun_out/ncu_u0.log 2>&1
echo "c128 k11 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:unit64_tc_kernel -s 27 -c 1 -o gpurun_out/prof_u64_c32k3 -f python tools/ncu_forward.py > gpurun_out/ncu_u1.log 2>&1
echo "c32 k3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:unit64_tc_kernel -s 24 -c 1 -o gpurun_out/prof_u64_c64k11 -f python tools/ncu_forward.py > gpurun_out/ncu_u2.log 2>&1
echo "c64 k11 rc=$?"
ls -la gpurun_out/*.ncu-rep
