"""torch.profiler view of one training step: CUDA time of this library's kernels vs everything torch launches around them."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wav2sleep_b200 import build_default  # noqa: E402
from wav2sleep_b200.optim import FusedAdamW  # noqa: E402
from wav2sleep_b200.trainer import SignalMasker, SleepLightningModule  # noqa: E402

dev = torch.device("cuda:0")
model = build_default(bench.CARDIO, 4, seed=0).to(dev)
masker = SignalMasker({"ABD": 0.7, "THX": 0.7, "ECG": 0.5, "PPG": 0.1}, backups=["ECG", "PPG"])
pl = SleepLightningModule(model, optimizer=lambda ps: FusedAdamW(ps, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0),
                          num_classes=4, masker=masker)
pl.setup_training()
torch.manual_seed(0)
src = {k: v.to(dev) for k, v in bench.make_night_batch(16, seed=7).items()}
y = torch.randint(0, 4, (16, bench.S_EPOCHS), device=dev)
for _ in range(3):
    pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
    torch.cuda.synchronize()
ours, other = 0.0, {}
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", 0) or getattr(ev, "cuda_time_total", 0)
    if ev.device_type.name != "CUDA" or t == 0:
        continue
    if "w2s" in ev.key:
        ours += t
    else:
        other[ev.key] = other.get(ev.key, 0) + t
print(f"library kernels {ours / 1e3:.1f} ms; other CUDA work {sum(other.values()) / 1e3:.1f} ms:")
for k, v in sorted(other.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {v / 1e3:7.2f} ms  {k[:110]}")
