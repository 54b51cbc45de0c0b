#!/bin/bash
# ncu --set full of every launch of the dominant kernel FUNCTION of the step (the 16-channel stride-2 stream kernel, 8
# launches per step) for profiles/traffic.json
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 500 $NCU --set full --kernel-name-base demangled -k 'regex:conv_stream_kernel<\(int\)16, \(int\)16, \(int\)2,' -c 8 \
   -o gpurun_out/r02_c16s2 python tools/profile_step.py infer 16 > gpurun_out/ncu_c16s2.log 2>&1
tail -1 gpurun_out/ncu_c16s2.log
ncu -i gpurun_out/r02_c16s2.ncu-rep --page raw --csv > gpurun_out/r02_c16s2.csv 2>/dev/null
rm -f gpurun_out/r02_c16s2.ncu-rep; ls -la gpurun_out/r02_c16s2.csv
