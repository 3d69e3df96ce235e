mkdir -p gpurun_out
for cfg in "0 1" "1 1" "0 0" "1 0"; do
  set -- $cfg
  echo "=== CLUSTER=$1 FUSE=$2"
  VTTS_TC_CLUSTER=$1 VTTS_TC_FUSE=$2 VTTS_PROFILE=1 timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "vtts-prof" | tail -80 | awk '{k=$2" "$4" "$6" "$7" "$8; t[k]+=$(NF-3); n[k]++} END{for(k in t) printf "%s  %.3f ms (%d)\n", k, t[k], n[k]}' | sort
done
