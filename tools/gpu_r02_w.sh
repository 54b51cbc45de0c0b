#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ab2.sh "ntw12 ntw14" 2
python - <<'PY'
import json
ks = {v: {k["kernel"]: k for k in json.load(open(f"gpurun_out/ab_{v}_2_kernels.json"))} for v in ("default", "ntw12", "ntw14")}
for name, k in ks["default"].items():
    if "pro2" in name and ("c16->32" in name or "c32->32" in name):
        print(f"{name:62s} " + "  ".join(f"{v} {ks[v][name]['avg_ms']*1e3:7.1f}" for v in ks))
PY
