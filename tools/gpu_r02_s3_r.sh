#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "paired or golden or full_night_argmax" > gpurun_out/r_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/r_tests.log
bash tools/gpu_env_ab.sh "W2S_PDL=1" 3
