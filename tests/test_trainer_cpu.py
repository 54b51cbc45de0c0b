"""CPU tests of the training harness host logic: masker semantics, LR schedule, bucketed gradient all-reduce (gloo x2)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wav2sleep_b200.optim import ExpWarmUpScheduler
from wav2sleep_b200.trainer import GradReducer, SignalMasker, confusion_matrix, invert_signals


def test_masker_keeps_at_least_one_signal_and_respects_backups():
    torch.manual_seed(0)
    masker = SignalMasker({"ABD": 0.7, "THX": 0.7, "ECG": 0.5, "PPG": 0.1}, backups=["ECG", "PPG"])
    drops = {k: 0 for k in ("ABD", "THX", "ECG", "PPG")}
    n = 0
    for _ in range(50):
        x = {k: torch.randn(16, 8) for k in drops}
        x["PPG"][3] = float("-inf")  # PPG missing for sample 3 before masking
        masker(x)
        miss = torch.stack([torch.isinf(v[:, 0]) for v in x.values()], -1)
        assert not miss.all(-1).any()          # every sample keeps >= 1 signal
        assert miss[3, 3]                       # an unavailable signal stays unavailable
        assert all(torch.isinf(v).all(-1).eq(torch.isinf(v[:, 0])).all() for v in x.values())  # whole rows
        for i, k in enumerate(drops):
            drops[k] += miss[:, i].sum().item()
        n += 16
    assert 0.6 < drops["ABD"] / n < 0.8 and 0.4 < drops["ECG"] / n < 0.6 and drops["PPG"] / n < 0.25
    with pytest.raises(ValueError):
        SignalMasker({"ECG": 0.5})({"ECG": torch.full((2, 4), float("-inf"))})
    with pytest.raises(ValueError):
        SignalMasker({"ECG": 1.5})
    with pytest.raises(ValueError):
        SignalMasker({"ECG": 1.0, "PPG": 1.0})({"ECG": torch.randn(2, 4), "PPG": torch.randn(2, 4)})
    with pytest.raises(ValueError):  # no backup channel available for a sample whose draws all failed
        SignalMasker({"ECG": 0.5, "ABD": 0.5}, backups=["ECG"])({"ECG": torch.full((2, 4), float("-inf")),
                                                                  "ABD": torch.randn(2, 4)})


def test_invert_signals_and_confusion_matrix():
    torch.manual_seed(1)
    x = {"ECG": torch.ones(64, 5)}
    invert_signals(x)
    assert set(x["ECG"][:, 0].tolist()) == {1.0, -1.0} and (x["ECG"].abs() == 1).all()
    logits = torch.tensor([[2.0, 0, 0], [0, 3.0, 0], [0, 0, 1.0], [5.0, 0, 0]])
    cm = confusion_matrix(logits, torch.tensor([0, 1, 1, -1]), 3)
    assert cm.tolist() == [[1, 0, 0], [0, 1, 1], [0, 0, 0]]


def test_exp_warmup_scheduler_matches_reference_formula():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sched = ExpWarmUpScheduler(opt, lr_max=1e-3, warmup_steps=10, tau=100.0)
    lrs = []
    for _ in range(30):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    # values printed by the reference's own ExpWarmUpScheduler (trainer/scheduler.py) for these settings: the LR of
    # optimizer step k (0-based) uses step = k + 1
    assert lrs[:3] == pytest.approx([1e-4, 2e-4, 3e-4], abs=1e-15) and lrs[9] == pytest.approx(1e-3, abs=1e-15)
    assert lrs[10] == pytest.approx(0.000990049833749168, abs=1e-15)
    for k, lr in enumerate(lrs):
        step = k + 1
        want = 1e-3 * step / 10 if step <= 10 else 1e-3 * math.exp(-(step - 10) / 100.0)
        assert abs(lr - want) < 1e-12, (step, lr, want)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    red = GradReducer(flat, {"encoders": (0, 6), "tail": (6, 10)}, use_stream=False)
    red("tail")       # fired first by the backward
    after_tail = flat.clone()
    red("encoders")
    red.wait()
    q.put((rank, after_tail, flat.clone()))
    dist.destroy_process_group()


def test_grad_reducer_buckets_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = torch.arange(10, dtype=torch.float32)
    for rank, after_tail, final in res:
        assert torch.equal(after_tail[6:], base[6:] * 3)            # tail bucket summed over ranks (1x + 2x)
        assert torch.equal(after_tail[:6], base[:6] * (rank + 1))   # encoder bucket still local
        assert torch.equal(final, base * 3)


def test_validation_and_test_steps_re_evaluate_signal_subsets():
    """reference trainer/main.py:188-226: unified models are re-evaluated on ECG, ECG+THX, PPG, PPG+THX per dataset;
    host-side control flow only, so a stub model that records the signal sets it was called with is enough."""
    from wav2sleep_b200.trainer import SleepLightningModule

    class Stub(torch.nn.Module):
        valid_signals = ["ABD", "THX", "ECG", "PPG"]

        def __init__(self):
            super().__init__()
            self.signal_encoders = [0, 1, 2, 3]  # len() > 1 -> unified
            self.calls = []

        def forward(self, x):
            self.calls.append(tuple(x.keys()))
            B, T = next(iter(x.values())).shape
            return torch.zeros(B, T // 4, 4)

        def predict(self, x):
            return self.forward(x).argmax(-1)

    stub = Stub()
    pl = SleepLightningModule(stub, num_classes=4, masker=None)
    pl.val_dataset_map = {0: "all", 1: "shhs", 2: "mesa", 3: "cfs"}
    pl.test_dataset_map = {0: "shhs", 1: "mesa"}
    batch = ({k: torch.zeros(2, 16) for k in ("ABD", "THX", "ECG", "PPG")}, torch.zeros(2, 4, dtype=torch.long))
    full = ("ABD", "THX", "ECG", "PPG")
    pl.validation_step(batch, dataloader_idx=0)
    assert stub.calls == [full]                                        # combined loader: no subsets
    stub.calls.clear(); pl.validation_step(batch, dataloader_idx=1)    # SHHS: ECG, ECG+THX (no PPG there)
    assert stub.calls == [full, ("ECG",), ("ECG", "THX")]
    stub.calls.clear(); pl.validation_step(batch, dataloader_idx=2)    # MESA: + PPG, PPG+THX
    assert stub.calls == [full, ("ECG",), ("ECG", "THX"), ("PPG",), ("PPG", "THX")]
    stub.calls.clear(); pl.validation_step(batch, dataloader_idx=3)    # CFS: ECG, PPG
    assert stub.calls == [full, ("ECG",), ("PPG",)]
    stub.calls.clear(); pl.test_step(batch, dataloader_idx=0)          # test: ECG+THX wherever THX exists
    assert stub.calls == [full, ("ECG",), ("ECG", "THX")]
    assert set(pl.aux_outputs["val"]) >= {(None, "all"), ("ECG", "shhs"), ("ECG_THX", "mesa"), ("PPG_THX", "mesa")}
    assert int(pl.aux_outputs["val"][("ECG", "mesa")].sum()) == 8 and int(pl.cmats["val"].sum()) == 4 * 8
    out = pl.predict_step(batch)
    assert set(out) == {"labels", "preds_ECG", "preds_ECG_THX", "preds"} and out["preds"].shape == (2, 4)


def test_lightning_module_with_a_unimodal_model():
    """reference trainer/main.py:97-114: for SleepPPGNet the module drops the masker, is not 'unified' and hands the
    model the single tensor of a one-key input dict."""
    from wav2sleep_b200.trainer import SignalMasker, SleepLightningModule

    class Uni(torch.nn.Module):  # no signal_encoders attribute, like SleepPPGNet
        valid_signals = ["PPG"]

        def forward(self, x_BT):
            assert isinstance(x_BT, torch.Tensor)
            return torch.zeros(x_BT.shape[0], x_BT.shape[1] // 4, 4)

    pl = SleepLightningModule(Uni(), num_classes=4, masker=SignalMasker({"PPG": 0.1}))
    assert pl.masker is None and not pl.unified
    batch = ({"PPG": torch.zeros(2, 16)}, torch.zeros(2, 4, dtype=torch.long))
    assert pl.validation_step(batch).ndim == 0 and ("PPG", "all") in pl.aux_outputs["val"]
    with pytest.raises(ValueError):
        pl({"PPG": torch.zeros(2, 16), "ECG": torch.zeros(2, 16)})
