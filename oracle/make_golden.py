"""Generate tests/golden/*.npz by running the REAL reference model (read-only at /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py

The reference package cannot be imported normally (its __init__ pulls hydra/lightning, absent here), so the
model files are imported through a bare package stub.  Weights are the reference's own
default initialisation under torch.manual_seed(SEED); they are NOT stored - the fixture stores a per-tensor
checksum so the tests can prove that ``wav2sleep_b200.build_default(seed=SEED)`` reproduces them bit-for-bit,
plus the reference logits / intermediate features for seeded inputs.
"""
from __future__ import annotations

import hashlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF_SRC = "/root/reference/src/wav2sleep"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
SEED = 0

CASES = {
    # name: (signal_map, num_classes, B, S, masked (signal, row) pairs, absent signals)
    "cardio_b2_s8": ({"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}, 4, 2, 8, [], []),
    "cardio_masked_b3_s8": ({"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}, 4, 3, 8,
                            [("ABD", 0), ("PPG", 0), ("ECG", 1), ("THX", 2)], []),
    "cardio_ecg_ppg_only_b2_s5": ({"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}, 4, 2, 5, [], ["ABD", "THX"]),
    "eog_b2_s4": ({"EOG-L": "EOG-L", "EOG-R": "EOG-R"}, 5, 2, 4, [("EOG-R", 1)], []),
}
SPE = {"ABD": 256, "THX": 256, "ECG": 1024, "PPG": 1024, "EOG-L": 4096, "EOG-R": 4096}


def import_reference():
    pkg = types.ModuleType("wav2sleep")
    pkg.__path__ = [REF_SRC]
    sys.modules["wav2sleep"] = pkg
    import wav2sleep.models.wav2sleep as m  # noqa: E402
    return m


def build_reference(m, signal_map, num_classes):
    """Literal kwargs of scripts/config/model/wav2sleep.yaml + main.yaml:21-22."""
    torch.manual_seed(SEED)
    enc = m.SignalEncoders(signal_map=signal_map, feature_dim=128, activation="gelu", norm="instance", causal=False,
                           chunk_causal=False, initial_channels=16, max_channels=128, output_norm=False,
                           use_residual=True)
    mix = m.MultiModalAttentionEmbedder(feature_dim=128, dropout=0.1, activation="gelu", layers=2, dim_ff=512, nhead=8)
    seq = m.SequenceCNN(feature_dim=128, dropout=0.1, activation="gelu", norm="layer", causal=False, num_layers=2,
                        kernel_size=7, num_dilations=6)
    return m.Wav2Sleep(enc, mix, seq, num_classes).eval()


def make_inputs(signal_map, B, S, masked, absent, seed=42):
    g = torch.Generator().manual_seed(seed)
    x = {}
    for name in signal_map:
        t = torch.randn(B, S * SPE[name], generator=g)
        if name in absent:
            continue
        x[name] = t
    for name, row in masked:
        x[name][row] = float("-inf")
    return x


def tensor_digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    m = import_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    for name, (smap, ncls, B, S, masked, absent) in CASES.items():
        model = build_reference(m, smap, ncls)
        sd = model.state_dict()
        x = make_inputs(smap, B, S, masked, absent)
        with torch.inference_mode():
            z = model.signal_encoders(x)
            mix = model.epoch_mixer(z)
            seq = model.sequence_mixer(mix)
            logits = model(x)
        out = {
            "logits": logits.numpy(), "mixer": mix.numpy(), "seq": seq.numpy(),
            "sd_keys": np.array(list(sd.keys())),
            "sd_digest": np.array([tensor_digest(v) for v in sd.values()]),
            "sd_shapes": np.array([str(tuple(v.shape)) for v in sd.values()]),
            "meta": np.array([str(dict(signal_map=smap, num_classes=ncls, B=B, S=S, masked=masked, absent=absent,
                                       seed=SEED, input_seed=42, torch=torch.__version__))]),
        }
        for sig, zz in z.items():
            out[f"z_{sig}"] = zz.numpy()
        np.savez_compressed(OUT / f"{name}.npz", **out)
        print(name, "logits", tuple(logits.shape), "params", sum(v.numel() for v in sd.values()))


if __name__ == "__main__":
    main()
