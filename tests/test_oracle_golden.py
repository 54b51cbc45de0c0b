"""CPU tests: the oracle restatement and the parameter mirror are pinned to the real reference.

Golden fixtures were produced by oracle/make_golden.py from the reference model itself (reference
models/wav2sleep.py:48-67 forward, default init under torch.manual_seed(0)).
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden, make_inputs
from oracle import wav2sleep_oracle as oracle
from wav2sleep_b200 import build_default

CASES = ["cardio_b2_s8", "cardio_masked_b3_s8", "cardio_ecg_ppg_only_b2_s5", "eog_b2_s4"]


def digest(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("case", CASES)
def test_mirror_init_matches_reference_state_dict(case):
    """Same keys, shapes and bit-identical default init as the reference (183 tensors for the cardio model)."""
    g, meta = load_golden(case)
    model = build_default(meta["signal_map"], meta["num_classes"], seed=meta["seed"])
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["sd_keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["sd_shapes"]]
    assert [digest(v) for v in sd.values()] == [str(d) for d in g["sd_digest"]]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_outputs(case):
    g, meta = load_golden(case)
    model = build_default(meta["signal_map"], meta["num_classes"], seed=meta["seed"])
    cfg = oracle.OracleConfig(signal_map=meta["signal_map"], num_classes=meta["num_classes"])
    x = make_inputs(meta["signal_map"], meta["B"], meta["S"], meta["masked"], meta["absent"], meta["input_seed"])
    logits, inter = oracle.forward(x, model.state_dict(), cfg, return_intermediates=True)
    for sig, z in inter["z"].items():
        ref = torch.from_numpy(g[f"z_{sig}"])
        assert torch.equal(torch.isinf(z), torch.isinf(ref))
        fin = ~torch.isinf(ref)
        assert (z[fin] - ref[fin]).abs().max().item() < 1e-4, sig  # fp32 op-order noise on O(10) values
    assert (inter["mixer"] - torch.from_numpy(g["mixer"])).abs().max().item() < 1e-4
    assert (inter["seq"] - torch.from_numpy(g["seq"])).abs().max().item() < 1e-4
    assert (logits - torch.from_numpy(g["logits"])).abs().max().item() < 5e-5
    assert logits.shape == (meta["B"], meta["S"], meta["num_classes"])


def test_oracle_masked_equals_absent():
    """SURVEY section 4 invariant: a signal passed as all -inf == the key being absent from the dict."""
    smap = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
    model = build_default(smap, 4, seed=0)
    cfg = oracle.cardio_config()
    x = make_inputs(smap, 2, 4)
    x_masked = {k: v.clone() for k, v in x.items()}
    x_masked["ABD"][:] = float("-inf")
    x_absent = {k: v for k, v in x.items() if k != "ABD"}
    a = oracle.forward(x_masked, model.state_dict(), cfg)
    b = oracle.forward(x_absent, model.state_dict(), cfg)
    assert (a - b).abs().max().item() < 1e-6


def test_oracle_errors():
    smap = {"ECG": "ECG"}
    model = build_default(smap, 4, seed=0)
    cfg = oracle.OracleConfig(signal_map=smap, num_classes=4)
    with pytest.raises(ValueError):
        oracle.forward({"ECG": torch.zeros(1, 1000)}, model.state_dict(), cfg)
    with pytest.raises(ValueError):
        oracle.epoch_mixer({}, model.state_dict(), cfg)


def test_oracle_training_graph_matches_reference_with_dropout_masks():
    """Train-mode pin (oracle/make_golden_train.py): the reference in train() with nn.Dropout patched to explicit keep
    masks -> logits, loss and parameter gradients; the oracle's training graph with the same masks must reproduce them."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "train_dropout_cardio.npz")
    smap = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
    B, S = g["labels"].shape
    model = build_default(smap, 4, seed=int(g["weights_seed"]))
    x = make_inputs(smap, B, S, masked=[("ABD", 0), ("PPG", 1)], seed=int(g["input_seed"]))
    masks = {}
    for i, name in enumerate(g["mask_names"]):
        shape = tuple(int(v) for v in g[f"mask_{i}_shape"])
        n = int(np.prod(shape))
        masks[str(name)] = torch.from_numpy(np.unpackbits(g[f"mask_{i}_bits"])[:n].reshape(shape).copy())
    pre = "epoch_mixer.transformer_encoder.layers"
    drop = {"p_mix": float(g["mask_p"][0]), "p_seq": float(g["mask_p"][-1]),
            "mix": [{"sa": masks[f"{pre}.{l}.dropout1"], "ff": masks[f"{pre}.{l}.dropout"], "out": masks[f"{pre}.{l}.dropout2"]}
                    for l in range(2)],
            "seq": [masks[f"sequence_mixer.dilated_convs.{b}.dropout"].transpose(1, 2) for b in range(2)]}  # [B,F,S]->[B,S,F]
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    logits = oracle.forward_with_grad(x, params, oracle.cardio_config(), dropout=drop)
    labels = torch.from_numpy(g["labels"])
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, 4), labels.reshape(-1), ignore_index=-1)
    loss.backward()
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < 1e-4
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    for name, ref_norm in zip(g["grad_norm_names"], g["grad_norms"]):
        gr = params[str(name)].grad
        got = 0.0 if gr is None else gr.norm().item()
        assert abs(got - ref_norm) <= 2e-3 * max(ref_norm, 1e-6) + 1e-7, (name, got, ref_norm)
    for key in g.files:
        if key.startswith("grad::"):
            ref = torch.from_numpy(g[key])
            got = params[key[6:]].grad
            assert (got - ref).norm().item() <= 2e-3 * ref.norm().item() + 1e-7, key
    # and the eval graph differs (the masks matter)
    with torch.no_grad():
        ev = oracle.forward_with_grad(x, params, oracle.cardio_config())
    assert (ev - logits).abs().max().item() > 1e-2
