#!/bin/bash
# ncu passes (one gpurun call): launch list + full captures of selected conv_tc launches.
# conv_tc launch order per forward: 0 pre, 1 up0, 2-19 stage0, 20 up1, 21-38 stage1, 39 up2, 40-57 stage2, 58 up3, 59-76 stage3
mkdir -p gpurun_out
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -1 gpurun_out/ncu_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_tc|conv_post|cf_to_cl|conv_fp32" -s 79 -c 80 --csv --log-file gpurun_out/launches.csv python tools/ncu_forward.py > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
# stage1 k=11 conv1 (idx 33), conv2 (34)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 110 -c 2 -o gpurun_out/prof_stage1 -f python tools/ncu_forward.py > gpurun_out/ncu_s1.log 2>&1
echo "stage1 rc=$?"
# stage3: k=3 conv1/conv2 (59,60), k=11 conv1/conv2 (71,72), last conv (76)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 136 -c 2 -o gpurun_out/prof_stage3_k3 -f python tools/ncu_forward.py > gpurun_out/ncu_s3a.log 2>&1
echo "stage3 k3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 148 -c 2 -o gpurun_out/prof_stage3_k11 -f python tools/ncu_forward.py > gpurun_out/ncu_s3b.log 2>&1
echo "stage3 k11 rc=$?"
ls -la gpurun_out/*.ncu-rep
