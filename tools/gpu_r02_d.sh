#!/bin/bash
# Round 2, GPU call D: tests + bench after the wait / weight-load / epilogue changes, one source-level capture.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -120 > gpurun_out/gpu_tests.log
grep -E "median rel|worst encoder|passed|failed|FAILED|Error|agreement" gpurun_out/gpu_tests.log | tail -30
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 10 --warmup 3 --kernels-out gpurun_out/kernels_full.json > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
echo "bench done $(( $(date +%s) - T0 )) s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_step.py infer 16 > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --clock-control none --profile-from-start off --set full --import-source on --kernel-name-base demangled \
  -k 'regex:conv_stream_kernel<\(int\)16, \(int\)16, \(int\)1, \(int\)3' \
  -c 1 -o gpurun_out/r02d_src python tools/profile_step.py infer 16 ECG > gpurun_out/ncu_src.log 2>&1
tail -1 gpurun_out/ncu_src.log
ncu -i gpurun_out/r02d_src.ncu-rep --page source --csv --print-source sass > gpurun_out/r02d_src_sass.csv 2>/dev/null
ncu -i gpurun_out/r02d_src.ncu-rep --page raw --csv > gpurun_out/r02d_src_raw.csv 2>/dev/null
rm -f gpurun_out/r02d_src.ncu-rep
du -sh gpurun_out
echo "all done $(( $(date +%s) - T0 )) s"
