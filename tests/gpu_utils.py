"""Helpers for the -m gpu tests: call single kernels through the C ABI with torch tensors as device memory."""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from wav2sleep_b200 import _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def pack_conv(w, taps_major=0, taps=None, split=0):
    lib = _lib.load()
    w = w.detach().float().contiguous()
    if taps_major:
        cout, cin = w.shape[0], w.shape[1] // taps
    else:
        cout, cin, taps = w.shape
    out = torch.empty(taps * cin * cout * (2 if split else 1), dtype=torch.float16, device=w.device)
    _lib.check(lib.w2s_pack_conv_weight(w.data_ptr(), cout, cin, taps, taps_major, split, out.data_ptr(), stream()))
    return out


def uses_split(cin, cout):
    return _lib.load().w2s_conv_uses_split(cin, cout)


def sums(y_h):
    """[B, L, C] fp16 -> [B, C, 2] fp64 (sum, sum of squares over L), as the producing kernel would emit."""
    y = y_h.double()
    return torch.stack([y.sum(1), (y * y).sum(1)], dim=-1).contiguous()


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def prologue_ref(y_h, r_h=None, eps=1e-2):
    """What the conv prologue computes from stored pre-norm fp16 values (InstanceNorm over L + GELU [+res, GELU])."""
    y = y_h.float()
    mu = y.mean(1, keepdim=True)
    var = (y * y).mean(1, keepdim=True) - mu * mu
    a = gelu((y - mu) / torch.sqrt(var.clamp_min(0) + eps))
    if r_h is not None:
        a = gelu(a + r_h.float())
    return a


def conv_ref(a_BLC, w, stride=1, pad=1, dil=1, split=0):
    """fp32 conv of fp16-rounded operands (the tensor core multiplies fp16 x fp16 exactly, accumulates in fp32).
    split: operands are carried as fp16 hi+lo pairs in the kernel, i.e. effectively unrounded."""
    a = (a_BLC if split else a_BLC.half().float()).transpose(1, 2)
    w = w if split else w.half().float()
    return F.conv1d(a, w, None, stride=stride, padding=pad, dilation=dil).transpose(1, 2).contiguous()


def run_conv(**kw):
    """Fill a w2s_conv_call from keyword tensors / ints and launch it."""
    lib = _lib.load()
    c = _lib.ConvCall()
    keep = []
    for k, v in kw.items():
        field = "in_" if k == "in" else k
        if isinstance(v, torch.Tensor):
            assert v.is_cuda and v.is_contiguous(), k
            keep.append(v)
            setattr(c, field, v.data_ptr())
        elif v is not None:
            setattr(c, field, v)
    _lib.check(lib.w2s_conv1d_fwd(C.byref(c), stream()))
    torch.cuda.synchronize()
