#!/bin/bash
# One gpurun call: kernel tests, forward parity, smoke, bench, ncu launch list.  Logs go to gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/kernels.log
tail -15 gpurun_out/kernels.log
timeout 1200 python -m pytest tests/test_forward_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/forward.log
grep -E "max-abs|agreement|passed|failed|FAILED|Error" gpurun_out/forward.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
fi
