import ast
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
SPE = {"ABD": 256, "THX": 256, "ECG": 1024, "PPG": 1024, "EOG-L": 4096, "EOG-R": 4096}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    g = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    meta = ast.literal_eval(str(g["meta"][0]))
    return g, meta


def make_inputs(signal_map, B, S, masked=(), absent=(), seed=42):
    """Same generator as oracle/make_golden.py: every signal of signal_map drawn in order, then masks applied."""
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


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
