#!/bin/bash
# warp-parallel live list: GPU suite, then A/B against the serial-scan build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/o_tests.log
bash tools/gpu_ab2.sh "serial" 3
