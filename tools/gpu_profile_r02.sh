#!/bin/bash
# Round-2 evidence pass (one GPU): ncu launch list (time + DRAM bytes) of one untrimmed forward at the bench shape and
# `--set full` captures of every kernel class.  One forward of tools/ncu_forward.py = 40 launches matching $K:
#   cf_to_cl, conv_tc x23 (pre, up0, 18 x stage 0, up1, up2, up3), unit_tc x9 (stage 1), chain_tc x6 (stages 2, 3), conv_post_cl
mkdir -p gpurun_out
python tools/ncu_forward.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -1 gpurun_out/ncu_plain.log
for w in lr gauss path; do python tools/ncu_side_kernels.py $w >> gpurun_out/ncu_plain.log 2>&1 || { echo "side $w failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }; done
K='regex:conv_tc|unit_tc|unit64_tc|chain_tc|conv_post|cf_to_cl'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 40 -c 40 --csv --log-file gpurun_out/launches.csv python tools/ncu_forward.py > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
full() {  # name, kernel regex, skip, count, driver...
  local name=$1 rx=$2 s=$3 c=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $s -c $c -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
}
full prof_chain_c32_k11 chain_tc_kernel 11 1 python tools/ncu_forward.py
full prof_chain_c32_k3 chain_tc_kernel 9 1 python tools/ncu_forward.py
full prof_chain_c64_k7 chain_tc_kernel 7 1 python tools/ncu_forward.py
full prof_unit_c128_k11 unit_tc_kernel 15 1 python tools/ncu_forward.py
full prof_conv_c256_k11 conv_tc_kernel 37 2 python tools/ncu_forward.py
full prof_up1_256_128 conv_tc_kernel 43 1 python tools/ncu_forward.py
full prof_up3_64_32 conv_tc_kernel 45 1 python tools/ncu_forward.py
full prof_conv_post conv_post_cl_kernel 1 1 python tools/ncu_forward.py
full prof_lr_gather lr_gather_kernel 1 1 python tools/ncu_side_kernels.py lr
full prof_gauss gauss_upsample_kernel 1 1 python tools/ncu_side_kernels.py gauss
full prof_path 'path_generate_kernel|path_expand_kernel' 2 2 python tools/ncu_side_kernels.py path
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
# gpurun copies back at most 64 MiB: summarise the reports HERE and keep only the two chain-kernel reports
VTTS_PROFILE_OUT=gpurun_out/profiles_r02 python tools/make_profiles.py r02 > gpurun_out/make_profiles.log 2>&1
echo "make_profiles rc=$?"
for f in gpurun_out/*.ncu-rep; do case $f in *prof_chain_c32_k11*|*prof_chain_c64_k7*) ;; *) rm -f $f;; esac; done
du -sh gpurun_out
