#!/bin/bash
# compute-sanitizer memcheck over one tiny forward (smoke) and one tiny training step
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "smoke memcheck rc=$?"; tail -4 gpurun_out/memcheck_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider -k "training_reduces_loss or fused_adamw" > gpurun_out/memcheck_train.log 2>&1
echo "train memcheck rc=$?"; tail -4 gpurun_out/memcheck_train.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/initcheck_smoke.log 2>&1
echo "smoke initcheck rc=$?"; grep -E "Uninitialized|ERROR SUMMARY|smoke ok" gpurun_out/initcheck_smoke.log | sort | uniq -c | head
