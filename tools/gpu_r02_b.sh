#!/bin/bash
# Round 2, GPU call B: -m gpu suite, bench, launch list, ncu evidence at B = 4 nights with a bounded section set
# (ncu saves / restores the kernels' multi-GB workspaces for every replay pass: --set full at B = 16 takes ~10 min).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu -s -p no:cacheprovider -x 2>&1 | tail -250 > gpurun_out/gpu_tests.log
grep -E "median rel|worst encoder|passed|failed|FAILED|Error|agreement" gpurun_out/gpu_tests.log | tail -30
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
echo "bench done $(( $(date +%s) - T0 )) s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_step.py infer 16 > gpurun_out/ncu_launches.log 2>&1
echo "launch list done $(( $(date +%s) - T0 )) s"
SEC="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
NCU="ncu --clock-control none --profile-from-start off"
timeout 400 $NCU $SEC -c 64 -o gpurun_out/r02_infer_b4 python tools/profile_step.py infer 4 ECG,ABD > gpurun_out/ncu_infer.log 2>&1
tail -2 gpurun_out/ncu_infer.log
echo "ncu infer done $(( $(date +%s) - T0 )) s"
timeout 300 $NCU --set full --import-source on --kernel-name-base demangled \
  -k 'regex:conv_stream_kernel<\(int\)128, \(int\)128, \(int\)1, \(int\)1,|conv_stream_kernel<\(int\)64, \(int\)64, \(int\)1, \(int\)1,|conv_igemm_kernel<\(int\)128, \(int\)128, \(int\)7, \(int\)4, \(int\)0, \(int\)2' \
  -c 3 -o gpurun_out/r02_src_kernels python tools/profile_step.py infer 4 ECG > gpurun_out/ncu_src.log 2>&1
tail -2 gpurun_out/ncu_src.log
echo "ncu src done $(( $(date +%s) - T0 )) s"
timeout 400 $NCU $SEC --kernel-name-base demangled \
  -k 'regex:gemm_tn_kernel|enc_norm_bwd|adamw_kernel|sumsq_kernel|conv_igemm_kernel<.*\(int\)5, |enc_act_bwd|first_conv' \
  -c 90 -o gpurun_out/r02_train_b4 python tools/profile_step.py train 4 ECG > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
echo "ncu train done $(( $(date +%s) - T0 )) s"
timeout 200 python tools/profile_train.py 16 > gpurun_out/train_profile.txt 2>&1
head -24 gpurun_out/train_profile.txt
rm -f gpurun_out/*.ncu-rep.tmp; du -sh gpurun_out; ls -la gpurun_out | head -30
echo "all done $(( $(date +%s) - T0 )) s"
