#!/bin/bash
# PDL with the trigger at the end of each CTA's work (build "late") vs no PDL
mkdir -p gpurun_out
for r in 1 2 3; do
  for v in nopdl late_pdl; do
    if [ $v == nopdl ]; then envs="W2S_PDL=0"; else envs="W2S_PDL=1 W2S_LIB_VARIANT=late"; fi
    env $envs timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline > gpurun_out/ab_${v}_$r.json 2> gpurun_out/ab_${v}_$r.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_${v}_$r.json')); print('$v round $r: %.3f ms/step e2e %.3f serial %.3f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['whole_step']['kernel_ms_per_step'], d['clocks']['sm_mhz']))"
  done
done
