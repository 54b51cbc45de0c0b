"""Per-kernel CUDA-event profile of one training step (same setup as bench.py's train timing)."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wav2sleep_b200 import _lib, build_default  # noqa: E402
from wav2sleep_b200.optim import FusedAdamW  # noqa: E402
from wav2sleep_b200.trainer import SignalMasker, SleepLightningModule  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = build_default(bench.CARDIO, 4, seed=0).to(dev)
masker = None if "--nomask" in sys.argv else SignalMasker({"ABD": 0.7, "THX": 0.7, "ECG": 0.5, "PPG": 0.1}, backups=["ECG", "PPG"])
pl = SleepLightningModule(model, optimizer=lambda ps: FusedAdamW(ps, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0),
                          num_classes=4, masker=masker)
pl.setup_training()
torch.manual_seed(0)
src = {k: v.to(dev) for k, v in bench.make_night_batch(B, seed=7).items()}
if "--only-ecg" in sys.argv:  # BASELINE config 5: the other three signals are missing for every night
    for k in src:
        if k != "ECG":
            src[k].fill_(float("-inf"))
y = torch.randint(0, 4, (B, bench.S_EPOCHS), device=dev)
for _ in range(2):
    pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
torch.cuda.synchronize()
t0 = time.perf_counter()
pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
lib.w2s_profile_enable(1)
pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
torch.cuda.synchronize()
recs = bench.collect_profile(lib)
lib.w2s_profile_enable(0)
agg = {}
for label, ms, by, fl in recs:
    key = label.split(" B")[0] if label.startswith(("conv", "gemm_tn", "first_conv")) else label
    a = agg.setdefault(key, [0.0, 0, 0.0])
    a[0] += ms; a[1] += 1; a[2] += by
tot = sum(a[0] for a in agg.values())
print(f"host enqueue {t_host*1e3:.1f} ms, step wall {t_all*1e3:.1f} ms, kernels {tot:.1f} ms, launches {len(recs)}")
for k, (ms, n, by) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{k:55s} n={n:4d} {ms:8.2f} ms {100*ms/tot:5.1f}%  {by/ms*1e-6 if ms>0 else 0:7.0f} GB/s")
