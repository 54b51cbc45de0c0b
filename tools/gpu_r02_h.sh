#!/bin/bash
# fused sequence mixer: stage test, forward parity, A/B bench against the layer-per-launch path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -s -p no:cacheprovider -k "seqmixer or seq_conv" 2>&1 | tail -30
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -s -p no:cacheprovider -x -k "not eog and not fp32_check" 2>&1 | tail -15
for v in 1 0; do
  W2S_SEQ_FUSED=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline > gpurun_out/seqfused_$v.json 2> gpurun_out/seqfused_$v.err
  tail -2 gpurun_out/seqfused_$v.err
  python - <<PY
import json
b=json.load(open("gpurun_out/seqfused_$v.json"))
print("W2S_SEQ_FUSED=$v ms_per_step", b["ms_per_step"], "e2e", b["e2e"]["ms_per_step"], b["clocks"])
PY
done
