#!/bin/bash
# weight-gradient side stream of the tail backward: training tests, then train-step A/B (alternating, same box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/p_tests.log
for r in 1 2; do
  for v in 1 0; do
    echo "== W2S_WGRAD_STREAM=$v round $r"
    W2S_WGRAD_STREAM=$v timeout 300 python tools/time_train_configs.py 2>&1 | grep "ms/step"
  done
done
