"""predict_on_folder for parquet nights: file reading / batching / CSV writing on CPU, the whole
pipeline against the oracle on GPU."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import staging_oracle
from oracle import wav2sleep_oracle as oracle
from wav2sleep_b200 import build_default, folder

SMAP = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}


def _write_night(path, S, cols, seed, labels=True, datetime_index=False):
    """A night as process_waveform_dataframe leaves it: one row per ECG sample, slower signals / labels NaN-padded."""
    g = np.random.default_rng(seed)
    n = S * 1024
    data = {}
    for c in cols:
        spe = folder.COLS_TO_SAMPLES_PER_EPOCH[c]
        v = np.full(n, np.nan, dtype=np.float32)
        v[: S * spe] = (g.standard_normal(S * spe) * 40 + 7).astype(np.float32)
        data[c] = v
    if labels:
        lab = np.full(n, np.nan, dtype=np.float32)
        lab[:S] = g.integers(0, 5, S)
        data["Stage"] = lab
    idx = pd.date_range("2024-01-01 22:00", periods=n, freq="29296875ns") if datetime_index else None
    path.parent.mkdir(parents=True, exist_ok=True)
    pd.DataFrame(data, index=idx).to_parquet(path)


def test_load_night_and_batching(tmp_path):
    _write_night(tmp_path / "a" / "n1.parquet", 6, ["ECG", "ABD"], 1)
    _write_night(tmp_path / "a" / "n2.parquet", 6, ["ECG", "ABD"], 2)
    _write_night(tmp_path / "b" / "n3.parquet", 4, ["ECG", "ABD"], 3, labels=False)
    files = folder.parquet_files(str(tmp_path))
    assert [f.split("/")[-1] for f in files] == ["n1.parquet", "n2.parquet", "n3.parquet"]
    sig, lab = folder.load_night(files[0], list(SMAP), 4, max_length_hours=None)
    assert set(sig) == {"ECG", "ABD"} and sig["ECG"].shape == (6 * 1024,) and sig["ABD"].shape == (6 * 256,)
    assert lab.shape == (6,) and set(lab.tolist()) <= {0.0, 1.0, 2.0, 3.0}  # 5 -> 4 class mapping
    batches = list(folder.iter_batches(files, list(SMAP), 4, batch_size=4, pin=False))
    assert [b[0] for b in batches] == [[0, 1], [2]]  # equal-length runs only
    assert batches[0][1]["ECG"].shape == (2, 6 * 1024) and (batches[1][2] == -1).all()
    with pytest.raises(ValueError):
        folder.load_night(files[0], ["EOG-L"], 4)  # no relevant columns
    with pytest.raises(ValueError):
        folder.load_night(files[0], ["XYZ"], 4)


def test_save_predictions_csv_format(tmp_path):
    _write_night(tmp_path / "in" / "x" / "n1.parquet", 3, ["ECG"], 1, datetime_index=True)
    _write_night(tmp_path / "in" / "n2.parquet", 2, ["ECG"], 2)
    files = folder.parquet_files(str(tmp_path / "in"))
    preds = [torch.tensor([1, 2]), torch.tensor([0, 3, 1])]
    folder.save_predictions(preds, files, str(tmp_path / "in"), str(tmp_path / "out"), ["ECG"],
                            labels=[torch.tensor([1.0, -1.0]), torch.tensor([0.0, 3.0, 2.0])])
    a = pd.read_csv(tmp_path / "out" / "n2.preds.csv", index_col=0)
    assert a.index.name == "Timestamp" and list(a.index) == [30.0, 60.0] and list(a["Pred"]) == [1, 2]
    assert list(a["Stage"]) == [1.0, -1.0]
    b = pd.read_csv(tmp_path / "out" / "x" / "n1.preds.csv", index_col=0, parse_dates=True)
    assert list(b["Pred"]) == [0, 3, 1] and str(b.index[0]) == "2024-01-01 22:00:30"
    with pytest.raises(NotImplementedError):
        folder.predict_on_folder(str(tmp_path / "in"), str(tmp_path / "o2"), model=object(), preprocess=True)


@pytest.mark.gpu
def test_predict_on_folder_matches_oracle_pipeline(cuda_device, tmp_path):
    S = 8
    _write_night(tmp_path / "in" / "n1.parquet", S, ["ECG", "PPG", "ABD", "THX"], 11)
    _write_night(tmp_path / "in" / "n2.parquet", S, ["ECG", "PPG", "ABD", "THX"], 12)
    _write_night(tmp_path / "in" / "sub" / "n3.parquet", S, ["ECG", "ABD"], 13, labels=False)  # PPG / THX missing -> -inf
    model = build_default(SMAP, 4, seed=0)
    sd = model.state_dict()
    preds, labels = folder.predict_on_folder(str(tmp_path / "in"), str(tmp_path / "out"), model=model,
                                             device=str(cuda_device), batch_size=2, max_length_hours=10,
                                             return_tensors=True)
    files = folder.parquet_files(str(tmp_path / "in"))
    assert len(preds) == 3 and labels is not None
    for i, fp in enumerate(files):
        sig, _ = folder.load_night(fp, list(SMAP), 4)
        x = {c: staging_oracle.stage(sig[c][None]) if c in sig
             else torch.full((1, S * folder.COLS_TO_SAMPLES_PER_EPOCH[c]), float("-inf")) for c in SMAP}
        ref = oracle.forward(x, sd, oracle.cardio_config())[0]
        srt = ref.sort(-1).values
        clear = (srt[:, -1] - srt[:, -2]) > 4e-2  # ignore near-ties of the random-init logits
        assert preds[i].shape == (S,)
        assert torch.equal(preds[i][clear], ref.argmax(-1)[clear])
    out = pd.read_csv(tmp_path / "out" / "sub" / "n3.preds.csv", index_col=0)
    assert list(out["Pred"]) == preds[2].tolist() and list(out.index) == [30.0 * (k + 1) for k in range(S)]

    # the reference's three-call form of the same pipeline (api.py:143-220): load_dataset -> predict -> save_predictions
    import wav2sleep_b200 as pkg
    ds = pkg.load_dataset(str(tmp_path / "in"), model.valid_signals, num_classes=4, max_length_hours=10)
    assert len(ds) == 3 and ds.files == files and set(ds[2][0]) == {"ECG", "ABD"}
    P, L = pkg.predict(model, ds, device=str(cuda_device), batch_size=2)
    assert P.shape == (3, S) and P.dtype == torch.int64 and L is not None and L.shape == (3, S)
    for i in range(3):
        assert torch.equal(P[i], preds[i])
    pkg.save_predictions(P, str(tmp_path / "in"), str(tmp_path / "out3"), ds, labels=L, overwrite=True)
    again = pd.read_csv(tmp_path / "out3" / "sub" / "n3.preds.csv", index_col=0)
    assert list(again["Pred"]) == preds[2].tolist() and list(again["Stage"]) == [-1.0] * S
