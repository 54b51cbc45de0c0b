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
        params = list(params)
        if params and isinstance(params[0], dict):
            if len(params) != 1:
                raise ValueError("FusedAdamW updates one flat buffer with one set of hyper-parameters: pass a single "
                                 "parameter group")
            params = list(params[0]["params"])
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
        self._params = params
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

    def add_param_group(self, param_group):
        if getattr(self, "param_groups", None):
            raise ValueError("FusedAdamW supports a single parameter group")
        super().add_param_group(param_group)

    @torch.no_grad()
    def _realias(self) -> None:
        """``step`` reads only the flat buffers, so ``p.data`` / ``p.grad`` must still be views of them.  A foreign
        ``model.zero_grad()`` (``p.grad = None``, after which autograd installs fresh gradient tensors), ``p.grad = t``
        or ``model.to(...)`` breaks the alias: gather what the user left in ``p.grad`` / ``p.data`` into the flat
        buffers and re-point the parameter (a missing gradient counts as zero).  Cost: 2 pointer compares per tensor."""
        gp, pp = self.flat_grad.data_ptr(), self.flat_param.data_ptr()
        for p in self._params:
            off, k = self.offsets[id(p)]
            if p.data_ptr() != pp + 4 * off or p.dtype != torch.float32:
                if p.device != self.flat_param.device:
                    raise RuntimeError("FusedAdamW: a parameter left the optimizer's device; rebuild the optimizer "
                                       "after model.to(...)")
                self.flat_param[off:off + k].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = self.flat_param[off:off + k].view(p.shape)
                self._bump()
            gr = p.grad
            if gr is None:
                self.flat_grad[off:off + k].zero_()
                p.grad = self.flat_grad[off:off + k].view(p.shape)
            elif gr.data_ptr() != gp + 4 * off or gr.dtype != torch.float32 or not gr.is_contiguous():
                self.flat_grad[off:off + k].copy_(gr.detach().reshape(-1).to(torch.float32))
                p.grad = self.flat_grad[off:off + k].view(p.shape)

    def grads_aliased(self) -> bool:
        """True when every ``p.grad`` is still its view of the flat gradient buffer (what the bucketed data-parallel
        all-reduce, which runs before ``step``, relies on)."""
        gp = self.flat_grad.data_ptr()
        return all(p.grad is not None and p.grad.data_ptr() == gp + 4 * self.offsets[id(p)][0] for p in self._params)

    # ---- checkpointing: what torch.optim.AdamW keeps in ``state`` (+ the EMACallback's state_dict) lives in flat buffers
    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {"step": self._step, "exp_avg": self.exp_avg.detach().clone(),
                       "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                       "ema": None if self.ema is None else self.ema.detach().clone(),
                       "ema_decay": self.ema_decay, "ema_start_step": self.ema_start_step,
                       "numel": self.flat_param.numel()}
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        if fused is None:
            raise ValueError("not a FusedAdamW state_dict (no 'fused' entry): the Adam moments would be reset silently")
        if fused["numel"] != self.flat_param.numel():
            raise ValueError(f"state_dict holds {fused['numel']} elements, the optimizer {self.flat_param.numel()}")
        self._step = int(fused["step"])
        self.exp_avg.copy_(fused["exp_avg"])
        self.exp_avg_sq.copy_(fused["exp_avg_sq"])
        self.ema_decay, self.ema_start_step = fused["ema_decay"], fused["ema_start_step"]
        if fused["ema"] is not None:
            if self.ema is None:
                self.ema = torch.empty_like(self.flat_param)
            self.ema.copy_(fused["ema"])
        else:
            self.ema = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        self._realias()
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
