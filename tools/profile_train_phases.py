"""Wall-clock phases of one training step (events on the main stream; synchronising between phases, so the sum is an
upper bound of the pipelined step): forward, loss, tail backward (classifier + mixers), encoder backward, optimizer."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wav2sleep_b200 import build_default  # noqa: E402
from wav2sleep_b200.optim import FusedAdamW  # noqa: E402
from wav2sleep_b200.trainer import SignalMasker, SleepLightningModule  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = build_default(bench.CARDIO, 4, seed=0).to(dev)
masker = SignalMasker({"ABD": 0.7, "THX": 0.7, "ECG": 0.5, "PPG": 0.1}, backups=["ECG", "PPG"])
pl = SleepLightningModule(model, optimizer=lambda ps: FusedAdamW(ps, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0),
                          num_classes=4, masker=masker)
pl.setup_training()
torch.manual_seed(0)
src = {k: v.to(dev) for k, v in bench.make_night_batch(B, seed=7).items()}
y = torch.randint(0, 4, (B, bench.S_EPOCHS), device=dev)
eng = model._get_engine()
marks = {}
orig_enc = eng._encoder_backward
first = {"seen": False}


def enc_bwd(*a, **k):
    if not first["seen"]:
        first["seen"] = True
        torch.cuda.synchronize()
        marks["tail_bwd_done"] = time.perf_counter()
    return orig_enc(*a, **k)


eng._encoder_backward = enc_bwd
for it in range(4):
    first["seen"] = False
    batch = pl.on_after_batch_transfer(({k: v.clone() for k, v in src.items()}, y), training=True)
    pl._opt.zero_grad()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.train()
    loss = pl.training_step(batch)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    pl._opt.step()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 2:
        print(f"forward+loss {1e3 * (t1 - t0):.1f} ms | backward {1e3 * (t2 - t1):.1f} ms (tail: classifier + mixers "
              f"{1e3 * (marks['tail_bwd_done'] - t1):.1f} ms, encoders {1e3 * (t2 - marks['tail_bwd_done']):.1f} ms) | "
              f"optimizer {1e3 * (t3 - t2):.1f} ms | total {1e3 * (t3 - t0):.1f} ms")
