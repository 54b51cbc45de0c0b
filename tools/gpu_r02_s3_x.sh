#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/x_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/x_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider -k "paired and (smap3 or smap0)" > gpurun_out/x_memcheck_paired.log 2>&1
echo "memcheck paired rc=$?"; tail -n 3 gpurun_out/x_memcheck_paired.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-eog --no-cpu-baseline > gpurun_out/x_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/x_bench.json')); print('bench %.3f ms/step e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
