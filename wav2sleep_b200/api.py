"""Inference API mirroring the in-scope part of the reference's ``wav2sleep.api`` (api.py:53-99, 163-190).

``load_model`` reads the reference's artefact format (``config.yaml`` + ``state_dict.pth``) without hydra: a
small recursive ``_target_`` instantiator maps the reference's class paths onto this package, so checkpoints and
configs written by the reference load unmodified.  ``predict`` runs batches through the CUDA forward;
``predict_sharded`` splits recordings over ranks (one process per GPU, no data-path collective) and gathers the
integer predictions.  File ingestion (EDF/parquet, ``prepare``/``save_predictions``) stays with the reference
(SURVEY section 8f, row N2).
"""
from __future__ import annotations

import importlib
import os
from typing import Iterable

import torch
import yaml

from . import model as _model

# reference class path -> class in this package
_TARGETS = {
    "wav2sleep.models.wav2sleep.Wav2Sleep": _model.Wav2Sleep,
    "wav2sleep.models.wav2sleep.SignalEncoders": _model.SignalEncoders,
    "wav2sleep.models.wav2sleep.MultiModalAttentionEmbedder": _model.MultiModalAttentionEmbedder,
    "wav2sleep.models.wav2sleep.SequenceCNN": _model.SequenceCNN,
    "wav2sleep.models.ppgnet.SleepPPGNet": "wav2sleep_b200.ppgnet.SleepPPGNet",  # scripts/config/model/ppgnet.yaml
}


def instantiate(cfg):
    """Minimal stand-in for ``hydra.utils.instantiate`` on a resolved (interpolation-free) config tree."""
    if isinstance(cfg, dict):
        built = {k: instantiate(v) for k, v in cfg.items() if k != "_target_"}
        if "_target_" not in cfg:
            return built
        target = cfg["_target_"]
        cls = _TARGETS.get(target)
        if isinstance(cls, str):
            mod, name = cls.rsplit(".", 1)
            cls = getattr(importlib.import_module(mod), name)
        if cls is None:
            if not target.startswith("wav2sleep_b200."):
                raise ValueError(f"_target_ {target!r} is not part of the wav2sleep model tree")
            mod, name = target.rsplit(".", 1)
            cls = getattr(importlib.import_module(mod), name)
        return cls(**built)
    if isinstance(cfg, list):
        return [instantiate(v) for v in cfg]
    return cfg


def _resolve_device(device: str) -> str:
    if device == "auto":
        device = "cuda" if torch.cuda.is_available() else "cpu"
    return device


def load_model(folder: str, device: str = "auto", compile: bool = False, revision: str | None = None,
               cache_dir: str | None = None):
    """Load ``config.yaml`` + ``state_dict.pth`` from a local folder (reference api.py:53-99, same signature).
    Hub download is out of scope: an ``hf://`` URI raises (fetch the two files with the reference's ``hub`` module
    first); ``revision`` / ``cache_dir`` only apply to hub URIs.  ``compile`` is accepted and ignored (there is no
    tracing compiler on this path, see ``Wav2Sleep.compile``)."""
    if str(folder).startswith("hf://"):
        raise NotImplementedError("wav2sleep_b200.load_model reads local checkpoint folders only; download "
                                  f"{folder} (config.yaml + state_dict.pth) first")
    config_fp = os.path.join(folder, "config.yaml")
    if not os.path.exists(config_fp):
        raise FileNotFoundError(f"No config file found at {config_fp}. Has the model been downloaded?")
    with open(config_fp) as f:
        cfg = yaml.safe_load(f)
    model = instantiate(cfg)
    ckpt_path = os.path.join(folder, "state_dict.pth")
    if not os.path.exists(ckpt_path):
        raise FileNotFoundError(f"No state dict found at {ckpt_path}. Has the model been downloaded?")
    model.load_state_dict(torch.load(ckpt_path, weights_only=True))
    model.eval()
    return model.to(_resolve_device(device))


def default_config(signal_map: dict, num_classes: int) -> dict:
    """The resolved config tree of scripts/config/model/wav2sleep.yaml (what the reference logs as config.yaml)."""
    return {
        "_target_": "wav2sleep.models.wav2sleep.Wav2Sleep", "num_classes": num_classes,
        "signal_encoders": {"_target_": "wav2sleep.models.wav2sleep.SignalEncoders", "signal_map": dict(signal_map),
                            "feature_dim": 128, "activation": "gelu", "norm": "instance", "causal": False,
                            "chunk_causal": False, "initial_channels": 16, "max_channels": 128, "output_norm": False,
                            "use_residual": True},
        "epoch_mixer": {"_target_": "wav2sleep.models.wav2sleep.MultiModalAttentionEmbedder", "feature_dim": 128,
                        "dropout": 0.1, "activation": "gelu", "layers": 2, "dim_ff": 512, "nhead": 8},
        "sequence_mixer": {"_target_": "wav2sleep.models.wav2sleep.SequenceCNN", "feature_dim": 128, "dropout": 0.1,
                           "activation": "gelu", "norm": "layer", "causal": False, "num_layers": 2, "kernel_size": 7,
                           "num_dilations": 6},
    }


@torch.inference_mode()
def predict(model, batches, device: str = "auto", batch_size: int = 4, num_workers: int = 4):
    """Reference ``predict`` (api.py:163-190).  With a dataset from ``load_dataset`` (``.files`` / ``.columns``): the
    reference's call and return value - ``(predictions int64 [N, S], labels or None)``.  With an iterable of input
    dicts: ``model(x).argmax(-1)`` per batch, returned as one int64 [N, S] tensor on the CPU (device -> host copies are
    queued per batch and synchronised once at the end)."""
    if hasattr(batches, "files") and hasattr(batches, "columns"):
        from .folder import predict_dataset
        return predict_dataset(model, batches, device=device, batch_size=batch_size, num_workers=num_workers)
    device = _resolve_device(device)
    outs, pending = [], []
    for x in batches:
        x = {k: v.to(device, non_blocking=True) for k, v in x.items()}
        pending.append((model.predict_async(x), x))  # two batches in flight (x kept alive until its result is awaited)
        if len(pending) > 2:
            outs.append(pending.pop(0)[0].wait())
    outs += [p.wait() for p, _ in pending]
    if not outs:
        return torch.empty(0, 0, dtype=torch.int64)
    return torch.cat(outs, dim=0).cpu()


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous shard of ``n_items`` recordings for ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def predict_sharded(predict_fn, n_items: int, epochs: int) -> torch.Tensor:
    """Run ``predict_fn(indices) -> int64 [len(indices), epochs]`` on this rank's shard and gather all shards in
    order on every rank.  The only collective is this gather of integer predictions (host side of SURVEY 8e)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return predict_fn(list(range(n_items)))
    rank, world = dist.get_rank(), dist.get_world_size()
    idx = list(shard_range(n_items, rank, world))
    local = predict_fn(idx) if idx else torch.empty(0, epochs, dtype=torch.int64)
    max_n = (n_items + world - 1) // world
    # NCCL moves device memory only: under an NCCL-only group the (CPU) predictions travel through this rank's GPU
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    pad = torch.full((max_n, epochs), -1, dtype=torch.int64, device=dev)
    pad[: local.size(0)] = local.to(dev)
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    parts = [g[: len(shard_range(n_items, r, world))] for r, g in enumerate(gathered)]
    return torch.cat(parts, dim=0).cpu()
