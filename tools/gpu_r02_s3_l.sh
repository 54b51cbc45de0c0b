#!/bin/bash
# round-2 session 3: direct-from-global transform (NR = 0) - parity of the default build (128-channel kernels) and of
# the all-kernels variant, then A/B against the ring build and the wider direct variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -m gpu -x -q > gpurun_out/l_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/l_tests.log
W2S_LIB_VARIANT=d16 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -m gpu -x -q > gpurun_out/l_tests_d16.log 2>&1
echo "tests d16 rc=$?"; tail -3 gpurun_out/l_tests_d16.log
bash tools/gpu_ab2.sh "ring d64 d32 d16" 2
python - <<'PY'
import json
for v in ("default", "ring", "d64", "d32", "d16"):
    ks = json.load(open(f"gpurun_out/ab_{v}_2_kernels.json"))
    for k in ks:
        if " k3 " in k["kernel"] and "pro3" not in k["kernel"] and "pro4" not in k["kernel"]:
            print(v, k["kernel"], f"{k['avg_ms']*1e3:.1f} us", f"{k.get('algo_GBps',0):.0f} GB/s", f"{k.get('algo_TFLOPs',0):.0f} TF")
PY
