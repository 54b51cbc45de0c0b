"""Forward time of the other BASELINE.json configs (parity-test cases, not bench lines): config 1 on GPU (cardio, one
10-h night, batch 1), config 2 (EOG model, 16 x 14-h nights), config 5 inference shape (ECG only through the 4-signal model)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from wav2sleep_b200 import build_default  # noqa: E402

dev = torch.device("cuda:0")
SPE = {"ABD": 256, "THX": 256, "ECG": 1024, "PPG": 1024, "EOG-L": 4096, "EOG-R": 4096}


def timeit(name, smap, ncls, B, S, hours, present=None, iters=10):
    model = build_default(smap, ncls, seed=0).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    x = {k: torch.randn(B, S * SPE[k], generator=g).to(dev) for k in smap}
    if present is not None:
        for k in x:
            if k not in present:
                x[k][:] = float("-inf")
    with torch.inference_mode():
        for _ in range(3):
            model.predict(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            model.predict(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: B={B} S={S}: {ms:.3f} ms/forward = {B * hours / ms * 1e3:.0f} recording-hours/s")
    del model, x
    torch.cuda.empty_cache()


CARDIO = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
timeit("config1 (cardio, 1 night)", CARDIO, 4, 1, 1200, 10.0, iters=50)
timeit("config3 (cardio, 16 nights)", CARDIO, 4, 16, 1200, 10.0)
timeit("config5-inference (ECG only via masks, 32 nights)", CARDIO, 4, 32, 1200, 10.0, present=["ECG"])
timeit("config2 (EOG, 16 x 14 h)", {"EOG-L": "EOG-L", "EOG-R": "EOG-R"}, 5, 16, 1680, 14.0, iters=5)
