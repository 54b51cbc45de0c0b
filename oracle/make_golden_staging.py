"""Generate tests/golden/zscore.npz with the reference's own ParquetDataset._zscore_normalize (build container only)."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference/src/wav2sleep"
for name, path in (("wav2sleep", REF), ("wav2sleep.data", REF + "/data")):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
import wav2sleep.data.dataset as ds  # noqa: E402

g = torch.Generator().manual_seed(11)
nights = {
    "gauss_offset": torch.randn(4096, generator=g) * 37.5 + 1200.0,        # ADC-like offset and gain
    "adc_int16": torch.randint(-2000, 2000, (2048,), generator=g).float(),
    "constant": torch.full((1024,), 3.25),                                   # std < eps -> divided by eps
    "tiny_std": torch.randn(1024, generator=g) * 1e-8 + 5.0,
    "with_nan": torch.cat([torch.randn(1020, generator=g), torch.tensor([float("nan")] * 4)]),  # passed through
}
out = ds.ParquetDataset._zscore_normalize(nights)
np.savez_compressed(Path(__file__).resolve().parent.parent / "tests" / "golden" / "zscore.npz",
                    **{f"in::{k}": v.numpy() for k, v in nights.items()}, **{f"out::{k}": v.numpy() for k, v in out.items()})
print({k: (float(v[:2].mean()) if torch.isfinite(v).all() else "nan") for k, v in out.items()})
