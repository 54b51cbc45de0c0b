#!/bin/bash
# Round 2, GPU call A: -m gpu suite (new full-scale gradient tests), bench, ncu evidence for every kernel class.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider -x 2>&1 | tail -150 > gpurun_out/gpu_tests.log
grep -E "median rel|worst encoder|passed|failed|FAILED|Error|agreement" gpurun_out/gpu_tests.log | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
W2S_ENC_STREAMS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_serial.json 2> gpurun_out/bench_serial.err
cut -c1-300 gpurun_out/bench_serial.json
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --set full -o gpurun_out/r02_infer_step python tools/profile_step.py infer 16 > gpurun_out/ncu_infer.log 2>&1
tail -2 gpurun_out/ncu_infer.log
timeout 600 $NCU --set full --import-source on --kernel-name-base demangled \
  -k 'regex:conv_stream_kernel<\(int\)128, \(int\)128, \(int\)1, \(int\)1,|conv_stream_kernel<\(int\)64, \(int\)64, \(int\)1, \(int\)1,|conv_igemm_kernel<\(int\)128, \(int\)128, \(int\)7, \(int\)4, \(int\)0, \(int\)2|conv_stream_kernel<\(int\)32, \(int\)32, \(int\)1, \(int\)1,' \
  -c 4 -o gpurun_out/r02_src_kernels python tools/profile_step.py infer 16 > gpurun_out/ncu_src.log 2>&1
tail -2 gpurun_out/ncu_src.log
timeout 900 $NCU --set full --kernel-name-base demangled \
  -k 'regex:gemm_tn_kernel|enc_norm_bwd|adamw_kernel|sumsq_kernel|conv_igemm_kernel<.*\(int\)5, |enc_act_bwd|first_conv' \
  -c 120 -o gpurun_out/r02_train_step python tools/profile_step.py train 16 ECG > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
timeout 300 python tools/profile_train.py 16 > gpurun_out/train_profile.txt 2>&1
head -30 gpurun_out/train_profile.txt
ls -la gpurun_out | head -40
