#!/bin/bash
# Full GPU pass: all gpu tests, smoke, default bench, reference arm, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
echo "bench reference rc=$?"; tail -1 gpurun_out/bench_reference.log
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_tc|conv_post|cf_to_cl|lr_" -s 79 -c 80 --csv --log-file gpurun_out/launches.csv python tools/ncu_forward.py > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 110 -c 2 -o gpurun_out/prof_stage1_k11 -f python tools/ncu_forward.py > gpurun_out/ncu_s1.log 2>&1
echo "stage1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 136 -c 2 -o gpurun_out/prof_stage3_k3 -f python tools/ncu_forward.py > gpurun_out/ncu_s3.log 2>&1
echo "stage3 rc=$?"
ls -la gpurun_out/*.ncu-rep
