#!/bin/bash
# A/B of library variants on the training step (same box, alternating):  tools/gpu_ab_train.sh "v1 v2" [rounds]
mkdir -p gpurun_out
VS="$1"; R=${2:-2}
for r in $(seq 1 $R); do
  for v in default $VS; do
    vv=$v; [ "$v" == "default" ] && vv=""
    W2S_LIB_VARIANT=$vv timeout 300 python tools/profile_train.py 16 > gpurun_out/abt_${v}_$r.txt 2>&1
    echo "$v round $r: $(head -1 gpurun_out/abt_${v}_$r.txt)"
  done
done
