#!/bin/bash
# x_stats and the encoder Linear in the paired launches too: GPU suite, memcheck of the paired test, A/B by switch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?"; tail -n 2 gpurun_out/w_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider -k "paired and (smap3 or smap1)" > gpurun_out/w_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 2 gpurun_out/w_memcheck.log
bash tools/gpu_env_ab.sh "W2S_PAIR_LINEAR=0 W2S_PAIR_XSTATS=0" 3
