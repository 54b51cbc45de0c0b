#!/bin/bash
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_training_gpu.py -m gpu -x -q -k "two_forwards" 2>&1 | tail -n 1; done
W2S_WGRAD_STREAM=0 timeout 300 python -m pytest tests/test_training_gpu.py -m gpu -x -q 2>&1 | tail -n 1
timeout 300 python -m pytest tests/test_training_gpu.py -m gpu -x -q 2>&1 | tail -n 1
