#!/bin/bash
# Round 2, GPU call C: ncu evidence, each report kept small (gpurun brings back at most 64 MiB).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -250 > gpurun_out/gpu_tests.log
grep -E "median rel|worst encoder|passed|failed|FAILED|Error|agreement" gpurun_out/gpu_tests.log | tail -30
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
echo "bench done $(( $(date +%s) - T0 )) s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_step.py infer 16 > gpurun_out/ncu_launches.log 2>&1
SEC="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
NCU="ncu --clock-control none --profile-from-start off"
timeout 400 $NCU $SEC -c 64 -o gpurun_out/r02_infer_b4 python tools/profile_step.py infer 4 ECG,ABD > gpurun_out/ncu_infer.log 2>&1
tail -1 gpurun_out/ncu_infer.log
ncu -i gpurun_out/r02_infer_b4.ncu-rep --page raw --csv > gpurun_out/r02_infer_b4.csv 2>/dev/null; rm -f gpurun_out/r02_infer_b4.ncu-rep
echo "ncu infer done $(( $(date +%s) - T0 )) s"
timeout 300 $NCU --set full --import-source on --kernel-name-base demangled \
  -k 'regex:conv_stream_kernel<\(int\)128, \(int\)128, \(int\)1, \(int\)1,|conv_stream_kernel<\(int\)16, \(int\)16, \(int\)1, \(int\)3' \
  -c 2 -o gpurun_out/r02_src_kernels python tools/profile_step.py infer 4 ECG > gpurun_out/ncu_src.log 2>&1
tail -1 gpurun_out/ncu_src.log
echo "ncu src done $(( $(date +%s) - T0 )) s"
timeout 400 $NCU $SEC --kernel-name-base demangled \
  -k 'regex:gemm_tn_kernel|enc_norm_bwd|adamw_kernel|sumsq_kernel|conv_igemm_kernel<.*\(int\)5, |enc_act_bwd|first_conv' \
  -c 90 -o gpurun_out/r02_train_b4 python tools/profile_step.py train 4 ECG > gpurun_out/ncu_train.log 2>&1
tail -1 gpurun_out/ncu_train.log
ncu -i gpurun_out/r02_train_b4.ncu-rep --page raw --csv > gpurun_out/r02_train_b4.csv 2>/dev/null; rm -f gpurun_out/r02_train_b4.ncu-rep
echo "ncu train done $(( $(date +%s) - T0 )) s"
du -sh gpurun_out; ls -la gpurun_out | head -30
