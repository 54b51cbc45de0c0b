#!/bin/bash
# Round 2, GPU call G: sequence-mixer night groups A/B (runtime switch), new tests, missing ncu captures, bench.
mkdir -p gpurun_out
T0=$(date +%s)
for r in 1 2; do for g in 1 4 8; do
  W2S_SEQ_GROUPS=$g timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline > gpurun_out/seq_g${g}_$r.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/seq_g${g}_$r.json')); print('seq groups $g round $r: %.3f ms/step e2e %.3f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz']))"
done; done
echo "A/B done $(( $(date +%s) - T0 )) s"
timeout 900 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/gpu_tests.log
grep -E "passed|failed|FAILED|Error|fp32 check|mixer max-abs" gpurun_out/gpu_tests.log | tail -20
echo "tests done $(( $(date +%s) - T0 )) s"
NCU="ncu --clock-control none --profile-from-start off --set full --kernel-name-base demangled"
timeout 300 $NCU -k 'regex:epoch_mixer|conv_igemm_kernel' -c 18 -o gpurun_out/r02_tail python tools/profile_step.py infer 16 > gpurun_out/ncu_tail.log 2>&1
ncu -i gpurun_out/r02_tail.ncu-rep --page raw --csv > gpurun_out/r02_tail.csv 2>/dev/null; rm -f gpurun_out/r02_tail.ncu-rep
timeout 300 $NCU -k 'regex:adamw_kernel|sumsq_kernel|attn_kernel|row_ln_bwd|row_ln_fwd|head_bwd|colsum|tokens' -c 14 -o gpurun_out/r02_train_small python tools/profile_step.py train 16 ECG > gpurun_out/ncu_train_small.log 2>&1
ncu -i gpurun_out/r02_train_small.ncu-rep --page raw --csv > gpurun_out/r02_train_small.csv 2>/dev/null; rm -f gpurun_out/r02_train_small.ncu-rep
echo "ncu done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 20 --warmup 5 --kernels-out gpurun_out/kernels_full.json > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
du -sh gpurun_out; echo "all done $(( $(date +%s) - T0 )) s"
