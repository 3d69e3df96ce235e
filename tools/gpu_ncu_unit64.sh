#!/bin/bash
# ncu full captures of two narrow-unit launches of the second forward (unit64 launch order: stage2 units 0-8, stage3 units 9-17)
mkdir -p gpurun_out
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -1 gpurun_out/ncu_plain.log
ncu --set full --clock-control none --import-source on -k regex:unit64_tc_kernel -s 27 -c 1 -o gpurun_out/prof_u64_c32k3 -f python tools/ncu_forward.py > gpurun_out/ncu_u1.log 2>&1
echo "c32 k3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:unit64_tc_kernel -s 33 -c 1 -o gpurun_out/prof_u64_c32k11 -f python tools/ncu_forward.py > gpurun_out/ncu_u2.log 2>&1
echo "c32 k11 rc=$?"
ls -la gpurun_out/*.ncu-rep
