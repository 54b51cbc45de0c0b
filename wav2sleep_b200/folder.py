"""Folder-level inference: ``predict_on_folder`` for preprocessed parquet nights.

Mirrors reference ``api.predict_on_folder`` / ``load_dataset`` / ``predict`` / ``save_predictions`` (api.py:142-301) and the
reading half of ``ParquetDataset.__getitem__`` (data/dataset.py:132-183) with three changes that the GPU path wants:
nights are read raw and normalised on the device (``staging.stage_batch``: whole-night z-score + ``-inf`` fill), batches
are prefetched by a reader thread into pinned memory while the GPU works, and predictions are copied back once at the end
instead of per batch.  Under ``torch.distributed`` the file list is sharded over ranks (one process per GPU, no data-path
collective; rank 0 writes the CSVs).  EDF/CSV ingestion (``prepare``) is out of scope: pass ``preprocess=False``.
"""
from __future__ import annotations

import logging
import os
import queue
import threading
from glob import glob
from pathlib import Path
from typing import Iterable, Optional

import numpy as np
import torch

from .api import _resolve_device, load_model, predict_sharded
from .model import COLS_TO_SAMPLES_PER_EPOCH
from .staging import stage_batch

logger = logging.getLogger(__name__)
LABEL, TIMESTAMP, PRED = "Stage", "Timestamp", "Pred"                       # settings.py:8-11
INTEGER_LABEL_MAPS = {4: {0: 0, 1: 1, 2: 1, 3: 2, 4: 3}, 5: {0: 0, 1: 1, 2: 2, 3: 3, 4: 4}}  # settings.py:53-56


def parquet_files(folder: str) -> list[str]:
    return sorted(glob(os.path.join(folder, "**/*.parquet"), recursive=True))  # api.py:314-315 (sorted: stable shards)


def load_night(fp: str, columns: list[str], num_classes: int, max_length_hours: Optional[int] = None):
    """One parquet night -> ({signal: raw fp32 [T_sig]}, labels fp32 [S]); the un-normalised half of
    ParquetDataset.__getitem__ (data/dataset.py:132-183): same column checks, truncation and label mapping."""
    import pandas as pd
    df = pd.read_parquet(fp)
    max_epochs = 1_000_000 if max_length_hours is None else max_length_hours * 60 * 2
    signals, epochs = {}, None
    for col in columns:
        if col not in COLS_TO_SAMPLES_PER_EPOCH:
            raise ValueError(f"Column {col} unrecognised.")
        if col not in df.columns:
            continue
        x = torch.from_numpy(df[col].dropna().values.astype(np.float32))
        if torch.isinf(x).any():
            raise ValueError(f"fp={fp!r} has inf. values for col={col!r}")
        ep = x.shape[0] // COLS_TO_SAMPLES_PER_EPOCH[col]
        if epochs is None:
            epochs = ep
        elif epochs != ep:
            raise ValueError(f"prev_inferred_recording_length_epochs={epochs} != inferred_recording_length_epochs={ep} for fp={fp!r}")
        signals[col] = x[: COLS_TO_SAMPLES_PER_EPOCH[col] * min(ep, max_epochs)]
    if not signals:
        raise ValueError(f"No relevant columns found in fp={fp!r}. self.columns={columns}")
    if LABEL in df.columns:
        lab = df[LABEL].dropna().map(INTEGER_LABEL_MAPS[num_classes])
        labels = torch.from_numpy(lab.fillna(-1).values.T.astype(np.float32))
        if labels.shape[0] != epochs:
            raise ValueError(f"labels.shape={tuple(labels.shape)} != inferred_recording_length_epochs={epochs} for fp={fp!r}")
        labels = labels[:max_epochs]
    else:
        labels = torch.full((min(epochs, max_epochs),), -1.0)
    return signals, labels


class NightDataset:
    """What the reference's ``load_dataset`` returns, as far as its callers use it (``ParquetDataset``: ``.files``,
    ``.columns``, ``len()``, ``dataset[i] -> (signal dict, labels)``, data/dataset.py:20-75, 132-186).  Items are the RAW
    night (fp32 per signal, labels fp32 [S]); the whole-night z-score and the ``-inf`` fill of missing signals happen on
    the device when ``predict`` stages a batch (staging.stage_batch) instead of per item on the CPU."""

    def __init__(self, parquet_fps: list[str], num_classes: int, columns: list[str],
                 max_length_hours: Optional[int] = None):
        if num_classes not in INTEGER_LABEL_MAPS:
            raise ValueError(f"Unsupported num_classes={num_classes}")
        self.files, self.num_classes, self.columns = list(parquet_fps), num_classes, list(columns)
        self.max_length_hours = max_length_hours

    def __len__(self) -> int:
        return len(self.files)

    def __getitem__(self, idx):
        return load_night(self.files[idx], self.columns, self.num_classes, self.max_length_hours)


def load_dataset(parquet_folder: str, signals: Iterable[str], num_classes: int = 4,
                 max_length_hours: Optional[int] = None) -> NightDataset:
    """reference api.py:143-160"""
    files = parquet_files(parquet_folder)
    if len(files) == 0:
        raise ValueError(f"No parquet files found in {parquet_folder}.")
    return NightDataset(files, num_classes, list(signals), max_length_hours)


def predict_dataset(model, dataset: NightDataset, device: str = "auto", batch_size: int = 4, num_workers: int = 4):
    """reference ``predict(model, dataset, device, batch_size, num_workers)`` (api.py:163-190): -> (predictions int64
    [N, S], labels [N, S] or None when no night has labels).  Nights of different lengths come back padded with -1
    (the reference's default collate would refuse them).  ``num_workers`` is accepted for compatibility: one reader
    thread prefetches into pinned memory."""
    device = _resolve_device(device)
    preds, labels = predict_files(model.to(device).eval(), dataset.files, dataset.columns, device, batch_size,
                                  dataset.max_length_hours)
    S = max((len(p) for p in preds), default=0)
    P = torch.full((len(preds), S), -1, dtype=torch.int64)
    L = torch.full((len(preds), S), -1.0)
    for i, (p_, l_) in enumerate(zip(preds, labels)):
        P[i, : len(p_)] = p_
        L[i, : len(l_)] = l_
    return P, (None if bool((L == -1).all()) else L)


def iter_batches(files: list[str], columns: list[str], num_classes: int, batch_size: int,
                 max_length_hours: Optional[int] = None, pin: bool = True, prefetch: int = 2):
    """Yields (indices, {signal: raw [b, T_sig]} pinned, labels [b, S]) for runs of consecutive files of equal length
    and equal present columns, read by a background thread ``prefetch`` batches ahead."""
    def batches():
        cur, key = [], None
        for i, fp in enumerate(files):
            sig, lab = load_night(fp, columns, num_classes, max_length_hours)
            k = (lab.shape[0], tuple(sorted(sig)))
            if cur and (k != key or len(cur) == batch_size):
                yield cur
                cur = []
            key = k
            cur.append((i, sig, lab))
        if cur:
            yield cur

    def collate(items):
        idx = [i for i, _, _ in items]
        x = {c: torch.stack([s[c] for _, s, _ in items]) for c in items[0][1]}
        if pin and torch.cuda.is_available():
            x = {c: t.pin_memory() for c, t in x.items()}
        return idx, x, torch.stack([l for _, _, l in items])

    q: queue.Queue = queue.Queue(maxsize=max(prefetch, 1))

    def worker():
        try:
            for items in batches():
                q.put(collate(items))
            q.put(None)
        except BaseException as e:  # surfaced in the consumer
            q.put(e)

    threading.Thread(target=worker, daemon=True).start()
    while True:
        item = q.get()
        if item is None:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


@torch.inference_mode()
def predict_files(model, files: list[str], signals: list[str], device, batch_size: int = 4,
                  max_length_hours: Optional[int] = None):
    """-> (list of int64 [S_i] predictions on the CPU, list of label tensors), one entry per file, in file order."""
    preds, labels, order = [], [], []
    for idx, raw, lab in iter_batches(files, signals, model.num_classes, batch_size, max_length_hours):
        x = stage_batch(raw, device, columns=signals)
        preds.append(model.predict(x))  # stays on the device: one synchronising copy at the end
        labels.extend(lab.unbind(0))
        order.extend(idx)
    flat = [p for batch in preds for p in batch.cpu().unbind(0)]
    out_p, out_l = [None] * len(files), [None] * len(files)
    for i, p, l in zip(order, flat, labels):
        out_p[i], out_l[i] = p, l
    return out_p, out_l


def save_predictions(predictions, files, parquet_folder, output_folder, columns=None, labels=None,
                     overwrite: bool = False, max_length_hours: Optional[int] = None) -> None:
    """CSV per night mirroring the input tree: index ``Timestamp`` (end of each 30-s epoch, or datetimes when the input
    has a DatetimeIndex), column ``Pred`` (+ ``Stage``).  Reference api.py:193-220.

    Two calling conventions: ``(predictions, files, parquet_folder, output_folder, columns, ...)`` and the reference's
    ``(predictions, parquet_folder, output_folder, dataset, labels=None, overwrite=False, max_length_hours=None)`` where
    ``dataset`` carries ``.files`` / ``.columns``.  Rows padded with -1 (ragged nights) are cut at the night's length."""
    import pandas as pd
    if isinstance(files, (str, os.PathLike)) and hasattr(output_folder, "files"):  # the reference's argument order
        dataset, ref_labels = output_folder, columns
        files, parquet_folder, output_folder, columns = dataset.files, files, parquet_folder, dataset.columns
        labels = ref_labels if ref_labels is not None else labels
    for idx, fp in enumerate(files):
        out_fp = str(Path(output_folder) / Path(fp).relative_to(parquet_folder).with_suffix(".preds.csv"))
        if os.path.exists(out_fp) and not overwrite:
            logger.warning(f"File {out_fp} exists. Skipping.")
            continue
        input_df = pd.read_parquet(fp)
        input_df = input_df[list(set(columns) & set(input_df.columns))]
        pred = predictions[idx]
        pred = pred[np.asarray(pred) >= 0] if len(pred) and int(np.asarray(pred).min()) < 0 else pred  # -1 padding
        n = int(len(pred))
        index = pd.Index(np.arange(0, 60 * n / 2, step=30) + 30.0, name=TIMESTAMP)
        if isinstance(input_df.index, pd.DatetimeIndex):
            index = input_df.index[0] + pd.to_timedelta(index, unit="s")
        out = pd.DataFrame({PRED: np.asarray(pred[:n])}, index=index)
        if labels is not None:
            out[LABEL] = np.asarray(labels[idx][:n])
        os.makedirs(os.path.dirname(out_fp), exist_ok=True)
        out.to_csv(out_fp)


def predict_on_folder(input_folder: str, output_folder: str, *, model=None, model_folder: Optional[str] = None,
                      signals: Optional[Iterable[str]] = None, device: str = "auto", batch_size: int = 4,
                      num_workers: int = 4, preprocess: bool = False, max_length_hours: int = 10, overwrite: bool = False,
                      compile: bool = False, return_tensors: bool = False):
    """Same signature as reference ``predict_on_folder`` (api.py:223-301).  ``num_workers`` is accepted for compatibility
    (a single reader thread feeds the GPU); ``preprocess=True`` (EDF/CSV ingestion) is not part of this package."""
    if preprocess:
        raise NotImplementedError("EDF/CSV preprocessing (`prepare`) is out of scope: preprocess with the reference and "
                                  "pass preprocess=False with a folder of parquet files")
    device = _resolve_device(device)
    if model is None:
        if model_folder is None:
            raise ValueError("Either `model` or `model_folder` must be provided.")
        model = load_model(model_folder, device=device, compile=compile)
    else:
        model = model.to(device)
        model.eval()
    if signals is None:
        signals = list(model.valid_signals)
    else:
        signals = list(signals)
        valid = set(model.valid_signals)
        if not set(signals).issubset(valid):
            raise ValueError(f"Invalid signal subset: {signals}. Valid signals are: {sorted(valid)}")
    files = parquet_files(input_folder)
    if len(files) == 0:
        raise ValueError(f"No parquet files found in {input_folder}.")

    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    cache: dict = {}

    def run(indices):  # this rank's shard -> padded [n, S_max] (-1 beyond a night's length)
        p, l = predict_files(model, [files[i] for i in indices], signals, device, batch_size, max_length_hours)
        cache.update({i: (pp, ll) for i, pp, ll in zip(indices, p, l)})
        S = max_epochs
        out = torch.full((len(indices), S), -1, dtype=torch.int64)
        for r, pp in enumerate(p):
            out[r, : len(pp)] = pp
        return out

    max_epochs = max_length_hours * 120
    padded = predict_sharded(run, len(files), max_epochs) if distributed else run(list(range(len(files))))
    preds = [padded[i][padded[i] >= 0] if i not in cache else cache[i][0] for i in range(len(files))]
    labels = None
    if not distributed:
        labs = [cache[i][1] for i in range(len(files))]
        if not all(bool((l == -1).all()) for l in labs):
            labels = labs
    if not distributed or dist.get_rank() == 0:
        save_predictions(preds, files, input_folder, output_folder, signals, labels=labels, overwrite=overwrite)
    return (preds, labels) if return_tensors else None
