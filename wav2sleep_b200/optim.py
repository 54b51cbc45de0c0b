"""Fused global-norm clipping + AdamW on flat fp32 buffers, and the reference's LR schedule.

Replaces ``clip_grad_norm_(1.0)`` + ``torch.optim.AdamW(lr=1e-3, weight_decay=1e-4)`` driven per step by Lightning
(reference scripts/config/training/main.yaml:21-22, training/optimizer/adamw.yaml, trainer/main.py:273-297) with two
kernel launches per step (``w2s_sumsq`` + ``w2s_adamw_step``), independent of the number of parameter tensors (183).
Parameters and gradients are re-pointed at views of two flat buffers, which is also what the data-parallel gradient
all-reduce operates on (one NCCL call per bucket, see trainer.py).
"""
from __future__ import annotations

import math

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LRScheduler

from . import _lib

WEIGHTS_EPOCH = 0  # bumped after every in-place parameter update so that engines re-pack their fp16 operand copies


class FusedAdamW(Optimizer):
    """``ema_decay``: also keep an exponential moving average of the parameters, updated inside the same kernel from step
    ``ema_start_step`` on (reference ``EMACallback(decay, start_step)``, trainer/callbacks.py:12-66); ``swap_to_ema`` /
    ``swap_to_original`` exchange it with the live weights around validation like the callback does."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None,
                 ema_decay: float | None = None, ema_start_step: int = 0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW runs on CUDA only (no CPU fallback)")
        n = sum(p.numel() for p in params)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.offsets = {}
        off = 0
        for p in params:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.detach().reshape(-1).to(torch.float32))
            p.data = self.flat_param[off:off + k].view(p.shape)
            p.grad = self.flat_grad[off:off + k].view(p.shape)
            self.offsets[id(p)] = (off, k)
            off += k
        self._step = 0
        if ema_decay is not None and not 0.0 <= ema_decay <= 1.0:
            raise ValueError(f"decay must be in [0, 1], got {ema_decay}")  # callbacks.py:33-34
        self.ema_decay, self.ema_start_step = ema_decay, ema_start_step
        self.ema = self.flat_param.clone() if ema_decay is not None else None  # callbacks.py:44-49
        self._swapped = None
        self.grad_scale = 1.0  # e.g. 1 / world_size after a SUM all-reduce
        self._bump()

    @staticmethod
    def _bump():
        global WEIGHTS_EPOCH
        WEIGHTS_EPOCH += 1

    def zero_grad(self, set_to_none: bool = False):  # gradients must stay views of the flat buffer
        self.flat_grad.zero_()

    def segment(self, params) -> tuple[int, int]:
        """[start, end) of the flat buffers covered by a contiguous run of parameters."""
        spans = [self.offsets[id(p)] for p in params]
        start = min(s for s, _ in spans)
        end = max(s + k for s, k in spans)
        return start, end

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        g = self.param_groups[0]
        self._step += 1
        st = torch.cuda.current_stream().cuda_stream
        n = self.flat_param.numel()
        clip = g["max_grad_norm"]
        gn_ptr = None
        if clip is not None and clip > 0:
            self.gnorm_sq.zero_()
            _lib.check(lib.w2s_sumsq(self.flat_grad.data_ptr(), n, self.gnorm_sq.data_ptr(), st))
            gn_ptr = self.gnorm_sq.data_ptr()
        b1, b2 = g["betas"]
        if self._swapped is not None:
            raise RuntimeError("step() while the EMA weights are swapped in")
        ema_on = self.ema is not None and self._step >= self.ema_start_step  # callbacks.py:51-53, 76-78
        _lib.check(lib.w2s_adamw_step(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), n, gn_ptr, float(g["lr"]), b1, b2, g["eps"],
                                      g["weight_decay"], float(clip or 0.0), float(self.grad_scale), self._step,
                                      self.ema.data_ptr() if ema_on else None, float(self.ema_decay or 0.0), st))
        self._bump()
        return loss

    @torch.no_grad()
    def swap_to_ema(self):
        """Evaluate with the averaged weights (callbacks.py:80-86); parameters stay views of the flat buffer."""
        if self.ema is None or self._swapped is not None:
            return
        self._swapped = self.flat_param.clone()
        self.flat_param.copy_(self.ema)
        self._bump()

    @torch.no_grad()
    def swap_to_original(self):
        if self._swapped is None:
            return
        self.flat_param.copy_(self._swapped)
        self._swapped = None
        self._bump()

    def grad_norm(self) -> float:
        """Global L2 norm of the (scaled) gradient seen by the last step (host sync; for logging)."""
        return math.sqrt(float(self.gnorm_sq.item())) * self.grad_scale


class ExpWarmUpScheduler(LRScheduler):
    """Linear warm-up to lr_max over warmup_steps, then exp(-(step - warmup)/tau)  (reference trainer/scheduler.py:7-32)."""

    def __init__(self, optimizer, lr_max: float, warmup_steps: int, tau: float):
        self.lr_max, self.warmup_steps, self.tau = lr_max, warmup_steps, tau
        self.num_param_groups = len(optimizer.param_groups)
        super().__init__(optimizer, last_epoch=-1)

    def get_lr(self):
        step = self.last_epoch + 1
        if step <= self.warmup_steps:
            lr = self.lr_max * (step / self.warmup_steps)
        else:
            lr = self.lr_max * math.exp(-(step - self.warmup_steps) / self.tau)
        return [lr] * self.num_param_groups
