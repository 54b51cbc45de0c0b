#!/bin/bash
# Round 2, GPU call E: tests (incl. general path) + bench.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -120 > gpurun_out/gpu_tests.log
grep -E "median rel|worst encoder|passed|failed|FAILED|Error|agreement|max-abs logit" gpurun_out/gpu_tests.log | tail -40
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 10 --warmup 3 --kernels-out gpurun_out/kernels_full.json > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
echo "bench done $(( $(date +%s) - T0 )) s"
W2S_LANES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-eog --no-cpu-baseline > gpurun_out/bench_lanes1.json 2> gpurun_out/bench_lanes1.err
cut -c1-300 gpurun_out/bench_lanes1.json
echo "all done $(( $(date +%s) - T0 )) s"
