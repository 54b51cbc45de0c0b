"""One inference step or one training step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    python tools/profile_step.py infer [B] [signals]  # cardio forward, B nights (default 16), encoders serialised;
                                                      # optional comma list of signals kept (the others are -inf rows)
    python tools/profile_step.py train [B] [signals]  # cardio training step, optional comma list of signals kept (others -inf)
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wav2sleep_b200 import build_default  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
rt = torch.cuda.cudart()
if mode == "infer":
    model = build_default(bench.CARDIO, 4, seed=0).to(dev).eval()
    model._get_engine().serial_groups = True  # the launches of a normal step (paired encoders), one stream after the other
    x = {k: v.to(dev) for k, v in bench.make_night_batch(B, seed=42).items()}
    if len(sys.argv) > 3:
        for k in x:
            if k not in sys.argv[3].split(","):
                x[k].fill_(float("-inf"))
    with torch.inference_mode():
        for _ in range(2):
            model.predict(x)
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        model.predict(x)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
else:
    from wav2sleep_b200.optim import FusedAdamW
    from wav2sleep_b200.trainer import SleepLightningModule
    keep = sys.argv[3].split(",") if len(sys.argv) > 3 else list(bench.CARDIO)
    model = build_default(bench.CARDIO, 4, seed=0).to(dev)
    pl = SleepLightningModule(model, optimizer=lambda ps: FusedAdamW(ps, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0),
                              num_classes=4, masker=None, flip_polarity=False)
    pl.setup_training()
    src = {k: v.to(dev) for k, v in bench.make_night_batch(B, seed=7).items()}
    for k in src:
        if k not in keep:
            src[k].fill_(float("-inf"))
    y = torch.randint(0, 4, (B, bench.S_EPOCHS), device=dev)
    for _ in range(2):
        pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    pl.fit_step(({k: v.clone() for k, v in src.items()}, y))
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
print("done", mode, B)
