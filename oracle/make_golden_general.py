"""Generate tests/golden/general_*.npz: logits of the REAL reference model (read-only at /root/reference, CPU) for
NON-DEFAULT configurations - what the general CUDA path (wav2sleep_b200/general.py) must match.

Run in the build container only:    python oracle/make_golden_general.py

Nothing but outputs is stored.  The weights are each model's own default initialisation under torch.manual_seed(seed);
``wav2sleep_b200.model`` reproduces it bit-for-bit (same constructors, same RNG order), which this script asserts
before writing and the tests re-check through the stored per-tensor SHA-256.  BatchNorm / GroupNorm / LayerNorm-style
parameters and running statistics (which initialise to constants) are then re-drawn by ``perturb`` - a deterministic
function of the state_dict that the tests apply to the mirror as well - so that those code paths are really exercised.
"""
from __future__ import annotations

import hashlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_SRC = "/root/reference/src/wav2sleep"
OUT = ROOT / "tests" / "golden"
SPE = {"ABD": 256, "THX": 256, "ECG": 1024, "PPG": 1024, "EOG-L": 4096, "EOG-R": 4096}

# name -> dict(seed, signal_map, num_classes, encoders kwargs, mixer kwargs, seq kwargs, B, S, masked rows)
CASES = {
    # the configuration of the reference's own tests/model/test_causality.py:11-39
    "general_causal_batch_relu": dict(
        seed=1, signal_map={"ECG": "ECG", "PPG": "PPG"}, num_classes=4,
        enc=dict(feature_dim=16, activation="relu", norm="batch", causal=True),
        mix=dict(feature_dim=16), seq=dict(feature_dim=16, causal=True, norm="batch"), B=2, S=6, masked=[("PPG", 1)]),
    # causal convolutions (no chunking), layer norm, SiLU
    "general_causalconv_layer_silu": dict(
        seed=2, signal_map={"ABD": "ABD", "ECG": "ECG"}, num_classes=5,
        enc=dict(feature_dim=32, activation="silu", norm="layer", causal=True, chunk_causal=False),
        mix=dict(feature_dim=32, layers=2, nhead=4, dim_ff=64, activation="silu", norm_first=False),
        seq=dict(feature_dim=32, causal=True, norm="layer", activation="silu", num_layers=1, num_dilations=3, kernel_size=5),
        B=2, S=5, masked=[]),
    # shared encoder + signal embeddings, register tokens, output norm, 'auto' norm, leaky ReLU
    "general_shared_embed_registers": dict(
        seed=3, signal_map={"ECG": "UNI", "PPG": "UNI", "ABD": "ABD"}, num_classes=3,
        enc=dict(feature_dim=64, activation="leaky", norm="auto", embed_signals=True, output_norm=True),
        mix=dict(feature_dim=64, layers=2, nhead=8, dim_ff=128, activation="leaky", register_tokens=2),
        seq=dict(feature_dim=64, norm="rms", activation="leaky", num_layers=2, num_dilations=4, kernel_size=3),
        B=3, S=4, masked=[("ABD", 0), ("ECG", 2)]),
    # group norm, no residual branch, narrow channels, GELU, no norm in the sequence mixer (conv bias)
    "general_group_nores_narrow": dict(
        seed=4, signal_map={"THX": "THX"}, num_classes=4,
        enc=dict(feature_dim=24, activation="gelu", norm="group", initial_channels=8, max_channels=32, use_residual=False),
        mix=dict(feature_dim=24, layers=1, nhead=2, dim_ff=48),
        seq=dict(feature_dim=24, norm=None, activation="gelu", num_layers=1, num_dilations=2, kernel_size=7),
        B=2, S=7, masked=[]),
}


def import_reference():
    pkg = types.ModuleType("wav2sleep")
    pkg.__path__ = [REF_SRC]
    sys.modules["wav2sleep"] = pkg
    import wav2sleep.models.wav2sleep as m
    return m


def build(mod, case):
    torch.manual_seed(case["seed"])
    enc = mod.SignalEncoders(signal_map=case["signal_map"], **case["enc"])
    mix = mod.MultiModalAttentionEmbedder(**case["mix"])
    seq = mod.SequenceCNN(**case["seq"])
    return mod.Wav2Sleep(enc, mix, seq, case["num_classes"]).eval()


def perturb(model, seed):
    """Re-draw every parameter / buffer that initialises to a constant (norm weights, biases, running statistics)."""
    g = torch.Generator().manual_seed(1000 + seed)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if "num_batches_tracked" in k or not v.dtype.is_floating_point:
                continue
            if k.endswith("running_var"):
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
            elif k.endswith("running_mean"):
                v.copy_(0.3 * torch.randn(v.shape, generator=g))
            elif v.numel() > 0 and (v == v.flatten()[0]).all():  # constant init: norm weights / biases
                v.copy_(v + 0.2 * torch.randn(v.shape, generator=g))


def make_inputs(case, seed=42):
    g = torch.Generator().manual_seed(seed)
    x = {n: torch.randn(case["B"], case["S"] * SPE[n], generator=g) for n in case["signal_map"]}
    for n, row in case["masked"]:
        x[n][row] = float("-inf")
    return x


def digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


PPG_SEED = 7


def build_ppgnet(cls):
    torch.manual_seed(PPG_SEED)
    return cls(n_classes=4, feature_dim=128).eval()


def ppg_input(seed=42):
    return torch.randn(1, 1228800, generator=torch.Generator().manual_seed(seed))


def main():
    ref = import_reference()
    from wav2sleep_b200 import model as mirror
    # SleepPPGNet baseline (models/ppgnet.py): one 10-h night
    import wav2sleep.models.ppgnet as ref_ppg
    from wav2sleep_b200 import ppgnet as mirror_ppg
    rm, mm = build_ppgnet(ref_ppg.SleepPPGNet), build_ppgnet(mirror_ppg.SleepPPGNet)
    assert list(rm.state_dict().keys()) == list(mm.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(rm.state_dict().values(), mm.state_dict().values()))
    perturb(rm, PPG_SEED)
    perturb(mm, PPG_SEED)
    assert digest(rm.state_dict()) == digest(mm.state_dict())
    with torch.no_grad():
        logits = rm(ppg_input())
    np.savez_compressed(OUT / "general_ppgnet.npz", logits=logits.numpy().astype(np.float32), sha=np.array([digest(rm.state_dict())]),
                        n_keys=np.array([len(rm.state_dict())]))
    print(f"general_ppgnet: {len(rm.state_dict())} tensors, logits {tuple(logits.shape)} abs-max {logits.abs().max():.3f}")
    for name, case in CASES.items():
        rm, mm = build(ref, case), build(mirror, case)
        sd_r, sd_m = rm.state_dict(), mm.state_dict()
        assert list(sd_r.keys()) == list(sd_m.keys()), (name, set(sd_r) ^ set(sd_m))
        for k in sd_r:
            assert sd_r[k].shape == sd_m[k].shape and torch.equal(sd_r[k], sd_m[k]), (name, k)
        perturb(rm, case["seed"])
        perturb(mm, case["seed"])
        assert digest(rm.state_dict()) == digest(mm.state_dict())
        x = make_inputs(case)
        with torch.no_grad():
            logits = rm(x)
        np.savez_compressed(OUT / f"{name}.npz", logits=logits.numpy(), sha=np.array([digest(rm.state_dict())]),
                            n_keys=np.array([len(sd_r)]))
        print(f"{name}: {len(sd_r)} tensors, logits {tuple(logits.shape)} abs-max {logits.abs().max():.3f}")


if __name__ == "__main__":
    main()
