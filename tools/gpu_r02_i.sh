#!/bin/bash
# same-box A/B: fused cluster sequence mixer vs layer-per-launch in 4 night groups (the previous default)
mkdir -p gpurun_out
for r in 1 2 3; do
  for v in fused old; do
    if [ $v == fused ]; then export W2S_SEQ_FUSED=1 W2S_SEQ_GROUPS=1; else export W2S_SEQ_FUSED=0 W2S_SEQ_GROUPS=4; fi
    timeout 300 python bench.py --steps 40 --warmup 5 --no-train --no-eog --no-cpu-baseline > gpurun_out/ab_seq_${v}_$r.json 2> gpurun_out/ab_seq_${v}_$r.err
    python - $v $r <<'PY'
import json, sys
v, r = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/ab_seq_{v}_{r}.json"))
print(f"{v:6s} round {r}: {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['ms_per_step']:.3f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  done
done
