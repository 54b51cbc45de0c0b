"""CPU oracle of the input staging step.  TEST INFRASTRUCTURE ONLY.

Restates ``ParquetDataset._zscore_normalize`` (reference data/dataset.py:76-87) for one night and the -inf fill of
missing signals (:170-173).  Pinned by tests/golden/zscore.npz, generated from the reference's own function by
oracle/make_golden_staging.py.
"""
import torch


def zscore_night(x: torch.Tensor) -> torch.Tensor:
    """1-D fp32 night -> (x - mean) / max(std, 1e-6) with the unbiased std; unchanged if empty or non-finite."""
    eps = 1e-6
    if x.numel() == 0 or not torch.isfinite(x).all():
        return x
    mu = x.mean()
    sd = x.std()  # unbiased (N - 1), torch default
    sd = sd if sd > eps else torch.tensor(eps, dtype=x.dtype)
    return (x - mu) / sd


def stage(raw_BT: torch.Tensor, present=None) -> torch.Tensor:
    raw = raw_BT.to(torch.float32)
    rows = []
    for b in range(raw.size(0)):
        if present is not None and not bool(present[b]):
            rows.append(torch.full_like(raw[b], float("-inf")))
        else:
            rows.append(zscore_night(raw[b]))
    return torch.stack(rows)
