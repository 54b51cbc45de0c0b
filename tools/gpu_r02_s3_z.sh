#!/bin/bash
# final state: smoke, GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/z_tests.log 2>&1
echo "tests rc=$?"; tail -n 2 gpurun_out/z_tests.log
timeout 600 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/z_bench.json')); print('%.3f ms/step value %.0f e2e %.0f roofline %.3f train %.1f ecg %.1f eog %.1f cpu %.1f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['train']['ms_per_step'], d['train_ecg_only']['ms_per_step'], d['eog']['ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches']))"
