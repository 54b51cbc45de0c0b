#!/bin/bash
# staged epilogue + 14 transform warps for the 32-channel conv1 kernels (build "st32") vs per-thread statistics + 10 warps
mkdir -p gpurun_out
W2S_LIB_VARIANT=st32 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -m gpu -x -q > gpurun_out/w_tests.log 2>&1
echo "tests st32 rc=$?"; tail -n 3 gpurun_out/w_tests.log
bash tools/gpu_ab2.sh "st32" 3
python - <<'PY'
import json
ks = {v: {k["kernel"]: k for k in json.load(open(f"gpurun_out/ab_{v}_2_kernels.json"))} for v in ("default", "st32")}
for name, k in ks["default"].items():
    if "pro2" in name and ("c16->32" in name or "c32->32" in name):
        print(f"{name:62s} default {k['avg_ms']*1e3:7.1f} us   st32 {ks['st32'][name]['avg_ms']*1e3:7.1f} us")
PY
