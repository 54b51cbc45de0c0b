"""Train-step timings of the BASELINE training configs (fwd + CE + bwd + clip + AdamW, dropout on):
config 4 = cardio 4-signal, 16 nights, config masker (what bench.py's `train` leg measures);
config 5 = ECG-only at batch 32: (a) the 4-signal model with PPG / ABD / THX rows all -inf, (b) the single-encoder
model of inputs/cardiorespiratory/ecg.yaml (ECG -> UNI)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wav2sleep_b200 import build_default  # noqa: E402
from wav2sleep_b200.optim import FusedAdamW  # noqa: E402
from wav2sleep_b200.trainer import SignalMasker, SleepLightningModule  # noqa: E402

dev = torch.device("cuda:0")


def run(name, smap, B, masker=None, mask_all_but=None):
    torch.manual_seed(0)
    model = build_default(smap, 4, seed=0).to(dev)
    pl = SleepLightningModule(model, optimizer=lambda ps: FusedAdamW(ps, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0),
                              num_classes=4, masker=masker)
    pl.setup_training()
    src = {k: v.to(dev) for k, v in bench.make_night_batch(B, seed=7).items() if k in smap}
    if mask_all_but is not None:
        for k in src:
            if k != mask_all_but:
                src[k].fill_(float("-inf"))
    y = torch.randint(0, 4, (B, bench.S_EPOCHS), device=dev)
    step = lambda: pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: B={B}: {ms:.1f} ms/step = {B * 10 / ms * 1e3:.0f} recording-hours/s trained, loss {float(loss):.3f}, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del pl, model
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


run("config4 (cardio, config masker)", bench.CARDIO, 16,
    masker=SignalMasker({"ABD": 0.7, "THX": 0.7, "ECG": 0.5, "PPG": 0.1}, backups=["ECG", "PPG"]))
run("config4 (cardio, no masking)", bench.CARDIO, 16)
run("config5a (4-signal model, only ECG present)", bench.CARDIO, 32, mask_all_but="ECG")
run("config5b (ECG -> UNI single-encoder model)", {"ECG": "UNI"}, 32)
