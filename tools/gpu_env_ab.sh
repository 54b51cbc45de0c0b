#!/bin/bash
# A/B of environment switches on one box, alternating rounds: tools/gpu_env_ab.sh "NAME=VAL NAME2=VAL2 ..." [rounds]
# (each entry is compared against the default environment)
mkdir -p gpurun_out
VS="$1"; R=${2:-2}
for r in $(seq 1 $R); do
  for v in default $VS; do
    tag=$(echo "$v" | tr '=' '_')
    if [ "$v" == "default" ]; then envs=""; else envs="$v"; fi
    env $envs timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline \
        --kernels-out gpurun_out/ab_${tag}_${r}_kernels.json > gpurun_out/ab_${tag}_$r.json 2> gpurun_out/ab_${tag}_$r.err
    python - "$tag" $r <<'PY'
import json, sys
v, r = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/ab_{v}_{r}.json"))
print(f"{v:16s} round {r}: {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['ms_per_step']:.3f}  serial kernels {d['roofline']['whole_step']['kernel_ms_per_step']:.3f} ms  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  done
done
