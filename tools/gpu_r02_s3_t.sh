#!/bin/bash
# Round 2, session 3 (second half of the evidence): forward tests, full bench line, ncu --set full of the dominant
# kernel (paired launches) and of the 64/128-channel kernels.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_tests.log 2>&1
echo "tests rc=$? $(( $(date +%s) - T0 )) s"; tail -n 3 gpurun_out/t_tests.log
timeout 600 python bench.py --kernels-out gpurun_out/t_kernels.json > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench rc=$? $(( $(date +%s) - T0 )) s"; tail -n 2 gpurun_out/t_bench.err; cut -c1-300 gpurun_out/t_bench.json
NCU="ncu --clock-control none --profile-from-start off"
timeout 400 $NCU --set full --kernel-name-base demangled -k 'regex:conv_stream_kernel<\(int\)16, \(int\)16, \(int\)2' -c 4 -o gpurun_out/t_c16s2 python tools/profile_step.py infer 16 > gpurun_out/t_ncu_c16s2.log 2>&1
tail -n 2 gpurun_out/t_ncu_c16s2.log
ncu -i gpurun_out/t_c16s2.ncu-rep --page raw --csv > gpurun_out/t_c16s2.csv 2>/dev/null; rm -f gpurun_out/t_c16s2.ncu-rep
timeout 400 $NCU --set full --kernel-name-base demangled -k 'regex:conv_stream_kernel<\(int\)128|conv_stream_kernel<\(int\)64' -c 18 -o gpurun_out/t_c128 python tools/profile_step.py infer 16 > gpurun_out/t_ncu_c128.log 2>&1
tail -n 2 gpurun_out/t_ncu_c128.log
ncu -i gpurun_out/t_c128.ncu-rep --page raw --csv > gpurun_out/t_c128.csv 2>/dev/null; rm -f gpurun_out/t_c128.ncu-rep
ls -la gpurun_out/t_*; echo "all done $(( $(date +%s) - T0 )) s"
