"""The reference algorithm through torch.compile(mode="max-autotune") on the same GPU (the reference's own best GPU
path: /root/reference/tests/model/test_compile.py:35, api.py:96-97), next to eager and to this repo, on the bench
workload.  Informational ("kernel to beat"); run by hand under a timeout - inductor's autotuning of
~100 conv layers takes minutes:

    timeout 600 python tools/time_compiled_reference.py [B]
"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import make_inputs  # noqa: E402
from oracle import wav2sleep_oracle as oracle  # noqa: E402
from wav2sleep_b200 import build_default  # noqa: E402

CARDIO = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
model = build_default(CARDIO, 4, seed=0).to(dev).eval()
sd = {k: v.detach() for k, v in model.state_dict().items()}
x = {k: v.to(dev) for k, v in make_inputs(CARDIO, B, 1200, seed=42).items()}
cfg = oracle.cardio_config()


def graph(x):
    z = oracle.signal_encoders(x, sd, cfg)
    m = oracle.epoch_mixer(z, sd, cfg)
    s = oracle.sequence_mixer(m, sd, cfg)
    return (s @ sd["classifier.weight"].t() + sd["classifier.bias"]).argmax(-1)


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_eager = timeit(lambda: graph(x), 3)
    print(f"eager (cuDNN / cuBLAS, fp32 + TF32 convs): {t_eager:.1f} ms = {B * 10 / t_eager * 1e3:.0f} recording-h/s", flush=True)
    with torch.inference_mode():
        t_ours = timeit(lambda: model.predict(x))
    print(f"this repo: {t_ours:.2f} ms = {B * 10 / t_ours * 1e3:.0f} recording-h/s", flush=True)
    for mode in ("default", "max-autotune-no-cudagraphs"):
        t0 = time.time()
        try:
            cg = torch.compile(graph, mode=None if mode == "default" else mode)
            cg(x)
            torch.cuda.synchronize()
            t_c = timeit(lambda: cg(x), 3)
            agree = (cg(x) == model.predict(x)).float().mean().item()
            print(f"torch.compile({mode}): {t_c:.1f} ms = {B * 10 / t_c * 1e3:.0f} recording-h/s (compile {time.time() - t0:.0f} s); "
                  f"this repo is {t_c / t_ours:.1f}x faster; argmax agreement {agree:.4f}", flush=True)
        except Exception as e:  # inductor / triton problems must not hide the numbers above
            print(f"torch.compile({mode}) failed after {time.time() - t0:.0f} s: {type(e).__name__}: {str(e)[:300]}", flush=True)
