#!/bin/bash
# same-box A/B: one lane vs two asynchronous lanes (the fused sequence mixer leaves ~68 SMs idle for 175 us per batch)
mkdir -p gpurun_out
for r in 1 2; do
  for v in 1 2; do
    W2S_LANES=$v timeout 300 python bench.py --steps 40 --warmup 6 --no-train --no-eog --no-cpu-baseline > gpurun_out/ab_lanes${v}_$r.json 2> gpurun_out/ab_lanes${v}_$r.err
    python - $v $r <<'PY'
import json, sys
v, r = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/ab_lanes{v}_{r}.json"))
print(f"lanes={v} round {r}: {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['ms_per_step']:.3f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  done
done
