#!/bin/bash
# ncu source-level (SASS) capture of the epoch mixer at B = 16
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 400 $NCU --set full --import-source on --kernel-name-base demangled -k 'regex:epoch_mixer_kernel' -c 1 \
   -o gpurun_out/r02_mixer_src python tools/profile_step.py infer 16 > gpurun_out/ncu_mixer_src.log 2>&1
tail -1 gpurun_out/ncu_mixer_src.log
ncu -i gpurun_out/r02_mixer_src.ncu-rep --page source --csv > gpurun_out/r02_mixer_source.csv 2>/dev/null
ncu -i gpurun_out/r02_mixer_src.ncu-rep --page raw --csv > gpurun_out/r02_mixer_raw.csv 2>/dev/null
ls -la gpurun_out/r02_mixer_*; rm -f gpurun_out/r02_mixer_src.ncu-rep
