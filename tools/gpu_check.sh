#!/bin/bash
# One gpurun call: descriptor probe, kernel tests, forward parity.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( cd tests/native && for v in "0 0" "0 3" "1 0" "1 3"; do timeout 60 ./umma_probe $v; echo "exit $?"; done ) > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/kernels.log
cat gpurun_out/kernels.log | tail -40
timeout 1200 python -m pytest tests/test_forward_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/forward.log
tail -50 gpurun_out/forward.log
