"""Input staging: whole-night z-score + -inf fill, reference data/dataset.py:76-87,170-173."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import staging_oracle

GOLD = Path(__file__).parent / "golden" / "zscore.npz"


def _cases():
    g = np.load(GOLD)
    return {k[4:]: (torch.from_numpy(g[k]), torch.from_numpy(g["out::" + k[4:]])) for k in g.files if k.startswith("in::")}


def test_oracle_zscore_matches_reference_golden():
    """The restatement against outputs of the reference's own ParquetDataset._zscore_normalize."""
    for name, (x, ref) in _cases().items():
        out = staging_oracle.zscore_night(x)
        if name == "with_nan":
            assert torch.equal(torch.isnan(out), torch.isnan(ref)) and torch.equal(out[:1020], ref[:1020]), name
        else:
            assert torch.equal(out, ref), name
    absent = staging_oracle.stage(torch.zeros(2, 8), present=[True, False])
    assert torch.isinf(absent[1]).all() and (absent[1] < 0).all() and torch.isfinite(absent[0]).all()


def test_stage_batch_rejects_bad_input_without_gpu():
    from wav2sleep_b200 import staging
    with pytest.raises(ValueError):
        staging.stage_batch({"XYZ": torch.zeros(1, 1024)}, "cpu")
    with pytest.raises(ValueError):
        staging.stage_batch({"ECG": torch.zeros(1, 1000)}, "cpu")
    with pytest.raises(RuntimeError):
        staging.zscore_on_device(torch.zeros(1, 1024))  # CPU tensor: no fallback


@pytest.mark.gpu
def test_zscore_kernel_matches_reference_golden(cuda_device):
    from wav2sleep_b200.staging import zscore_on_device
    for name, (x, ref) in _cases().items():
        out = zscore_on_device(x[None].to(cuda_device)).cpu()[0]
        if name == "with_nan":  # non-finite night: passed through unchanged
            assert torch.equal(torch.isnan(out), torch.isnan(ref)) and torch.equal(out[:1020], ref[:1020])
        else:
            # the reference sums in fp32, the kernel in fp64: agreement to a few fp32 ulps of the z-scored values
            assert (out - ref).abs().max().item() < 2e-5, (name, (out - ref).abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.int16])
def test_zscore_kernel_batch_dtypes_and_absent_rows(cuda_device, dtype):
    from wav2sleep_b200.staging import zscore_on_device
    g = torch.Generator().manual_seed(3)
    B, T = 5, 307200
    raw = (torch.randn(B, T, generator=g) * torch.tensor([1.0, 40.0, 300.0, 0.5, 7.0])[:, None] * 10
           + torch.tensor([0.0, 100.0, -800.0, 3.0, 50.0])[:, None])
    raw = raw.round().clamp(-30000, 30000).to(dtype) if dtype == torch.int16 else raw.to(dtype)
    present = torch.tensor([True, True, False, True, True])
    out = zscore_on_device(raw.to(cuda_device), present).cpu()
    ref = staging_oracle.stage(raw, present)
    assert torch.isinf(out[2]).all() and (out[2] < 0).all()
    keep = [0, 1, 3, 4]
    assert (out[keep] - ref[keep]).abs().max().item() < 5e-5
    assert abs(out[0].mean().item()) < 1e-4 and abs(out[0].std().item() - 1) < 1e-4


@pytest.mark.gpu
def test_staged_int16_batch_through_the_model(cuda_device):
    """Raw int16 nights -> stage_batch -> forward equals the oracle pipeline (CPU z-score -> oracle forward)."""
    from oracle import wav2sleep_oracle as oracle
    from wav2sleep_b200 import build_default
    from wav2sleep_b200.staging import stage_batch
    smap = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
    g = torch.Generator().manual_seed(5)
    B, S = 2, 10
    raw = {"ECG": (torch.randn(B, S * 1024, generator=g) * 300 + 50).round().to(torch.int16),
           "ABD": (torch.randn(B, S * 256, generator=g) * 90 - 400).round().to(torch.int16)}
    model = build_default(smap, 4, seed=0)
    x_ref = {k: staging_oracle.stage(v) for k, v in raw.items()}
    x_ref.update({k: torch.full((B, S * (1024 if k == "PPG" else 256)), float("-inf")) for k in ("PPG", "THX")})
    ref = oracle.forward(x_ref, model.state_dict(), oracle.cardio_config())
    model = model.to(cuda_device).eval()
    x = stage_batch({k: v.pin_memory() for k, v in raw.items()}, cuda_device, columns=list(smap))
    assert set(x) == set(smap) and torch.isinf(x["PPG"]).all()
    out = model(x).float().cpu()
    assert (out - ref).abs().max().item() < 2e-2
