#!/bin/bash
# Round 2, session 3: evidence for the final state - GPU suite, PDL A/B on top of the paired launches, full bench line,
# ncu launch list of one step (paired launches), ncu --set full of the dominant kernel and of the 128-channel kernels.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1
echo "tests rc=$? $(( $(date +%s) - T0 )) s"; tail -n 3 gpurun_out/s_tests.log
bash tools/gpu_env_ab.sh "W2S_PDL=1" 2
echo "A/B done $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --kernels-out gpurun_out/s_kernels.json > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$? $(( $(date +%s) - T0 )) s"; tail -n 2 gpurun_out/s_bench.err; cut -c1-400 gpurun_out/s_bench.json
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/s_launches.csv python tools/profile_step.py infer 16 > gpurun_out/s_ncu_launches.log 2>&1
tail -n 1 gpurun_out/s_ncu_launches.log
timeout 400 $NCU --set full --kernel-name-base demangled -k 'regex:conv_stream_kernel<16, 16, 2' -c 4 -o gpurun_out/s_c16s2 python tools/profile_step.py infer 16 > gpurun_out/s_ncu_c16s2.log 2>&1
ncu -i gpurun_out/s_c16s2.ncu-rep --page raw --csv > gpurun_out/s_c16s2.csv 2>/dev/null; rm -f gpurun_out/s_c16s2.ncu-rep
timeout 400 $NCU --set full --kernel-name-base demangled -k 'regex:conv_stream_kernel<128, 128|conv_stream_kernel<64, 128|conv_stream_kernel<64, 64' -c 12 -o gpurun_out/s_c128 python tools/profile_step.py infer 16 > gpurun_out/s_ncu_c128.log 2>&1
ncu -i gpurun_out/s_c128.ncu-rep --page raw --csv > gpurun_out/s_c128.csv 2>/dev/null; rm -f gpurun_out/s_c128.ncu-rep
ls -la gpurun_out/s_*; echo "all done $(( $(date +%s) - T0 )) s"
