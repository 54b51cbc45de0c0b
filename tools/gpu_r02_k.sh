#!/bin/bash
# Round 2, ncu evidence after the fused sequence mixer: launch list of one step, section tables of every inference
# kernel at B = 16 for the tail kernels (epoch mixer, fused sequence mixer), full capture of the fused sequence mixer.
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python tools/profile_step.py infer 16 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
timeout 400 $NCU --set full --import-source on --kernel-name-base demangled -k 'regex:seq_mixer_kernel|epoch_mixer_kernel' -c 2 \
   -o gpurun_out/r02_tail_fused python tools/profile_step.py infer 16 > gpurun_out/ncu_tail_fused.log 2>&1
tail -1 gpurun_out/ncu_tail_fused.log
ncu -i gpurun_out/r02_tail_fused.ncu-rep --page raw --csv > gpurun_out/r02_tail_fused.csv 2>/dev/null
ncu -i gpurun_out/r02_tail_fused.ncu-rep --page source --csv -k regex:seq_mixer > gpurun_out/r02_seq_fused_source.csv 2>/dev/null
ls -la gpurun_out/r02_tail_fused* gpurun_out/r02_seq_fused_source.csv
rm -f gpurun_out/r02_tail_fused.ncu-rep
