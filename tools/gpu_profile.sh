#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stream -s 2 -c 1 -o gpurun_out/prof_conv_c16_fir \
   python tools/profile_conv.py --fir --iters 2 > gpurun_out/ncu_c16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:epoch_mixer -c 1 -o gpurun_out/prof_mixer \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_mixer.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:conv_igemm_kernel.*7 -c 1 -o gpurun_out/prof_seqconv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_seq.log 2>&1
ls -la gpurun_out
