#!/bin/bash
# 2 GPUs: multi-GPU tests, then the bench line at N = 2 (NUMA binding of the ranks, paired launches, train all-reduce)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/v_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/v_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/v_bench2.json 2> gpurun_out/v_bench2.err
echo "bench rc=$?"; tail -n 3 gpurun_out/v_bench2.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/v_bench2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "e2e", "config")})
print(d["train"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
