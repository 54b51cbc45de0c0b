#!/bin/bash
# One gpurun call: the whole -m gpu suite, smoke, bench (+ optional ncu launch list).  Logs go to gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/gpu_tests.log
grep -E "max-abs|agreement|median rel|passed|failed|FAILED|Error" gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
fi
