"""CPU tests of the drop-in boundary: the C-ABI library loads and exports what include/w2s_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from wav2sleep_b200 import _lib, build_default
from wav2sleep_b200 import model as M

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    text = (ROOT / "include" / "w2s_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2s_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/w2s_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.w2s_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.w2s_conv1d_fwd(None, None) != 0
    assert b"null" in lib.w2s_last_error()
    assert lib.w2s_pack_conv_weight(None, 16, 12, 3, 0, 0, None, None) != 0
    assert b"pack_conv" in lib.w2s_last_error()


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout: compile a tiny probe with gcc against the header."""
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "w2s_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(w2s_conv_call), sizeof(w2s_encoder_desc), sizeof(w2s_mixer_layer), sizeof(w2s_mixer_desc), sizeof(w2s_seq_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "p.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", f"{d}/p", f"{d}/p.c"], check=True)
        out = subprocess.run([f"{d}/p"], capture_output=True, text=True, check=True).stdout.split()
    got = [ctypes.sizeof(c) for c in (_lib.ConvCall, _lib.EncoderDesc, _lib.MixerLayer, _lib.MixerDesc, _lib.SeqDesc)]
    assert got == [int(v) for v in out]


def test_model_surface_matches_reference_contract():
    smap = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
    m = build_default(smap, 4, seed=0)
    assert m.valid_signals == list(smap)
    assert m.num_classes == 4 and len(m.signal_encoders) == 4 and m.signal_encoders.causal is False
    assert sum(p.numel() for p in m.parameters()) == 2_948_740
    assert len(m.state_dict()) == 183
    # shared encoder (inputs/cardiorespiratory/ecg.yaml: ECG -> UNI)
    uni = build_default({"ECG": "UNI"}, 4, seed=0)
    assert list(uni.signal_encoders.encoders.keys()) == ["UNI"]


def test_reference_error_conventions():
    with pytest.raises(ValueError):
        M.SignalEncoders(signal_map={"EEG": "EEG"}, feature_dim=128, activation="gelu")
    with pytest.raises(ValueError):
        M.SignalEncoder(feature_dim=128, samples_per_epoch=1000)
    with pytest.raises(ValueError):  # models/utils.py:73-74
        M.SignalEncoders(signal_map={"ECG": "ECG"}, feature_dim=128, activation="tanh")
    with pytest.raises(NotImplementedError):  # the one reference option without a CUDA path
        M.SequenceCNN(norm="weight")
    # non-default options construct (general fp32 kernels) and are routed away from the fused path
    enc = M.SignalEncoders(signal_map={"ECG": "ECG"}, feature_dim=16, activation="relu", norm="batch", causal=True)
    model = M.Wav2Sleep(enc, M.MultiModalAttentionEmbedder(feature_dim=16), M.SequenceCNN(feature_dim=16, causal=True), 4)
    assert not model.fast_path and build_default({"ECG": "ECG"}, 4).fast_path


def test_cpu_input_fails_loudly():
    """No CPU fallback: a CPU tensor must raise, never silently run somewhere else."""
    m = build_default({"ECG": "ECG"}, 4, seed=0).eval()
    with pytest.raises(RuntimeError):
        m({"ECG": torch.zeros(1, 1024)})
    with pytest.raises(ValueError):
        m({})
