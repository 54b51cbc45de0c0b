#!/bin/bash
# paired encoder launches: GPU suite, then A/B W2S_ENC_PAIRS=0/1 (same build) and against the pre-pairing build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/q_tests.log
bash tools/gpu_env_ab.sh "W2S_ENC_PAIRS=0 W2S_LIB_VARIANT=prepair" 3
python - <<'PY'
import json
ks = {v: {k["kernel"]: k for k in json.load(open(f"gpurun_out/ab_{v}_2_kernels.json"))} for v in ("default", "W2S_ENC_PAIRS_0")}
tot = {v: sum(k["avg_ms"] * k["launches_per_step"] for k in ks[v].values()) for v in ks}
print("serial kernel ms per step:", tot)
for name, k in sorted(ks["default"].items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches_per_step"])[:40]:
    base = name.replace(" x2", "")
    o = ks["W2S_ENC_PAIRS_0"].get(base)
    print(f"{name:62s} n={k['launches_per_step']} {k['avg_ms']*1e3:7.1f} us" + (f"   single: n={o['launches_per_step']} {o['avg_ms']*1e3:7.1f} us" if o else ""))
PY
