#!/bin/bash
# round-2 session 3: programmatic dependent launch of the stream kernels - full GPU suite, then A/B against W2S_PDL=0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/m_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/m_tests.log
bash tools/gpu_env_ab.sh "W2S_PDL=0" 3
