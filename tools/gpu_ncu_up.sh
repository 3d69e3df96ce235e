#!/bin/bash
# ncu full captures of the upsample launches (conv_tc_kernel launch order per forward: pre, up0, 18 x stage0, up1, up2, up3)
mkdir -p gpurun_out
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -1 gpurun_out/ncu_plain.log
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 43 -c 3 -o gpurun_out/prof_ups -f python tools/ncu_forward.py > gpurun_out/ncu_up.log 2>&1
echo "ups rc=$?"
