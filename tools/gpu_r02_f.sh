#!/bin/bash
# Round 2, GPU call F: full -m gpu suite, then the torch.compile / eager baselines of the reference graph.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/gpu_tests.log
grep -E "median rel|worst encoder|passed|failed|FAILED|Error|agreement|max-abs logit|eager" gpurun_out/gpu_tests.log | tail -40
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 500 python tools/time_compiled_reference.py 16 > gpurun_out/compiled_reference.txt 2>&1
grep -E "eager|this repo|torch.compile" gpurun_out/compiled_reference.txt
echo "all done $(( $(date +%s) - T0 )) s"
