#!/bin/bash
# A/B on ONE box (boxes differ by a few % under the power cap): default library vs a variant, alternating runs.
# usage: tools/gpu_ab.sh <variant> [rounds]
mkdir -p gpurun_out
V=$1; R=${2:-2}
for r in $(seq 1 $R); do
  for v in "" "$V"; do
    W2S_LIB_VARIANT=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline \
        > gpurun_out/ab_${v:-default}_$r.json 2> gpurun_out/ab_${v:-default}_$r.err
    python - "$v" $r <<'PY'
import json, sys
v, r = sys.argv[1] or "default", sys.argv[2]
d = json.load(open(f"gpurun_out/ab_{v}_{r}.json"))
print(f"{v:12s} round {r}: {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['ms_per_step']:.3f}  serial kernels {d['roofline']['whole_step']['kernel_ms_per_step']:.3f} ms  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  done
done
