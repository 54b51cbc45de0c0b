#!/bin/bash
# A/B/C on one box: default vs several variants, R alternating rounds.   tools/gpu_ab2.sh "v1 v2 ..." [rounds]
mkdir -p gpurun_out
VS="$1"; R=${2:-2}
for r in $(seq 1 $R); do
  for v in default $VS; do
    vv=$v; [ "$v" == "default" ] && vv=""
    W2S_LIB_VARIANT=$vv timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline \
        --kernels-out gpurun_out/ab_${v}_${r}_kernels.json > gpurun_out/ab_${v}_$r.json 2> gpurun_out/ab_${v}_$r.err
    python - "$v" $r <<'PY'
import json, sys
v, r = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/ab_{v}_{r}.json"))
print(f"{v:12s} round {r}: {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['ms_per_step']:.3f}  serial kernels {d['roofline']['whole_step']['kernel_ms_per_step']:.3f} ms  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  done
done
