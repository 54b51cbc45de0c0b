#!/bin/bash
# 8 GPUs: the bench line at N = 8 (replicas, NUMA-bound ranks, paired launches; train all-reduce over NVLink)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/y_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/y_bench8.json 2> gpurun_out/y_bench8.err
echo "bench rc=$?"; tail -n 3 gpurun_out/y_bench8.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/y_bench8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "e2e")}, d["config"]["host_affinity"])
print({k: d["train"][k] for k in ("ms_per_step", "per_rank_ms", "same_mask_on_all_ranks")})
PY
