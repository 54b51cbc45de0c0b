"""fp32 check mode of the forward (north-star gate: logits within 1e-4 of the reference).

``forward_fp32(model, x)`` runs the whole model through the straightforward fp32 CUDA-core kernels of
``csrc/check_fp32.cuh`` (fp32 storage, exact erf GELU, fp64 statistics, PyTorch weight layouts).  It is slow (no
tensor cores, nothing fused) and exists only to validate the algorithm independently of 16-bit storage effects; the
product path never calls it.  torch is used for allocation and pure data movement (stack / index), never arithmetic.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib


def _st():
    return torch.cuda.current_stream().cuda_stream


def _f(t: Tensor) -> Tensor:
    return t.detach().to(torch.float32).contiguous()


class _Chk:
    def __init__(self, device):
        self.lib = _lib.load()
        self.dev = device

    def conv(self, x, w, B, L_in, L_out, cin, cout, taps, stride=1, dil=1, pad=0, mode=0, stats=None, res=None, bias=None,
             add=None, mask=None, taps_major=0, gelu_out=0, eps=0.0):
        out = torch.zeros(B, L_out, cout, dtype=torch.float32, device=self.dev)
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(self.lib.w2s_chk_conv(x.data_ptr(), p(res), p(stats), w.data_ptr(), p(bias), p(add), out.data_ptr(), p(mask),
                                         B, L_in, L_out, cin, cout, taps, stride, dil, pad, mode, taps_major, gelu_out, eps,
                                         _st()))
        return out

    def stats(self, y, mask):
        B, L, Cc = y.shape
        s = torch.zeros(B, Cc, 2, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.w2s_chk_stats(y.data_ptr(), s.data_ptr(), mask.data_ptr(), B, L, Cc, _st()))
        return s

    def rowln(self, x, g, b, gelu, eps, res=None):
        out = torch.empty_like(x)
        rows = x.numel() // 128
        _lib.check(self.lib.w2s_chk_rowln(x.data_ptr(), None if res is None else res.data_ptr(), g.data_ptr(), b.data_ptr(),
                                          out.data_ptr(), rows, gelu, eps, _st()))
        return out

    def attn(self, q, k, v, key_mask, N, D):
        o = torch.empty_like(q)
        _lib.check(self.lib.w2s_chk_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), key_mask.data_ptr(), N, D,
                                         _st()))
        return o


@torch.no_grad()
def forward_fp32(model, x: dict[str, Tensor]) -> Tensor:
    """dict of [B, T_sig] fp32 CUDA tensors -> logits [B, S, C]; same contract as ``Wav2Sleep.forward``."""
    eng = model._get_engine()
    B, S, device = eng._check_inputs(x)
    K = _Chk(device)
    names = sorted(x.keys())
    N = B * S
    with torch.cuda.device(device):
        feats, masks = [], []
        for n in names:
            enc = model.signal_encoders.get_encoder(n)
            xs = _f(x[n])
            T = xs.size(1)
            mask = torch.isinf(xs[:, 0]).to(torch.uint8).contiguous()  # data-dependent control, no arithmetic
            eps = enc.norm_eps
            L, cin = T, 1
            y3 = r = s3 = None
            for i, blk in enumerate(enc.cnn):
                c = enc.channels[i]
                w1, w2, w3, wd = (_f(blk.conv1.conv.weight), _f(blk.conv2.conv.weight), _f(blk.conv3.conv.weight),
                                  _f(blk.downsample.weight))
                if i == 0:
                    src = xs.view(B, T, 1)
                    y1 = K.conv(src, w1, B, L, L, 1, c, 3, pad=1, mode=3, mask=mask)
                    r_new = K.conv(src, wd, B, L, L // 2, 1, c, 1, stride=2, mode=3, mask=mask)
                else:
                    y1 = K.conv(y3, w1, B, L, L, cin, c, 3, pad=1, mode=2, stats=s3, res=r, mask=mask, eps=eps)
                    r_new = K.conv(y3, wd, B, L, L // 2, cin, c, 1, stride=2, mode=2, stats=s3, res=r, mask=mask, eps=eps)
                s1 = K.stats(y1, mask)
                y2 = K.conv(y1, w2, B, L, L, c, c, 3, pad=1, mode=1, stats=s1, mask=mask, eps=eps)
                s2 = K.stats(y2, mask)
                y3 = K.conv(y2, w3, B, L, L // 2, c, c, 3, stride=2, pad=1, mode=1, stats=s2, mask=mask, eps=eps)
                s3 = K.stats(y3, mask)
                r, cin, L = r_new, c, L // 2
            z = K.conv(y3, _f(enc.linear.weight), B, L, L // 4, cin, 128, 4, stride=4, mode=2, stats=s3, res=r,
                       bias=_f(enc.linear.bias), mask=mask, taps_major=1, gelu_out=1, eps=eps)
            feats.append(z)
            masks.append(mask)
        # ---- epoch mixer ----
        mix = model.epoch_mixer
        D = len(names) + 1
        cls = _f(mix.register_tokens[0, 0, :, 0])
        toks = [cls.expand(B, S, 128)] + [torch.where(m.bool()[:, None, None], torch.zeros_like(z), z)
                                          for z, m in zip(feats, masks)]
        xt = torch.stack(toks, dim=2).reshape(N * D, 128).contiguous()
        km = torch.stack([torch.zeros_like(masks[0])] + masks, dim=1)[:, None, :].expand(B, S, D).reshape(N * D).contiguous()
        T_tok = N * D
        for layer in mix.transformer_encoder.layers:
            eps = layer.norm1.eps
            h = K.rowln(xt, _f(layer.norm1.weight), _f(layer.norm1.bias), 0, eps)
            Win, bin_ = _f(layer.self_attn.in_proj_weight), _f(layer.self_attn.in_proj_bias)
            q, k, v = (K.conv(h, Win[j * 128:(j + 1) * 128].contiguous(), 1, T_tok, T_tok, 128, 128, 1,
                              bias=bin_[j * 128:(j + 1) * 128].contiguous()).view(T_tok, 128) for j in range(3))
            ao = K.attn(q, k, v, km, N, D)
            xt = K.conv(ao, _f(layer.self_attn.out_proj.weight), 1, T_tok, T_tok, 128, 128, 1,
                        bias=_f(layer.self_attn.out_proj.bias), add=xt).view(T_tok, 128)
            h = K.rowln(xt, _f(layer.norm2.weight), _f(layer.norm2.bias), 0, eps)
            hid = torch.cat([K.conv(h, _f(layer.linear1.weight)[j * 128:(j + 1) * 128].contiguous(), 1, T_tok, T_tok, 128,
                                    128, 1, bias=_f(layer.linear1.bias)[j * 128:(j + 1) * 128].contiguous(),
                                    gelu_out=1).view(T_tok, 128) for j in range(4)], dim=1).contiguous()
            xt = K.conv(hid, _f(layer.linear2.weight), 1, T_tok, T_tok, 512, 128, 1, bias=_f(layer.linear2.bias),
                        add=xt).view(T_tok, 128)
        cur = xt.view(N, D, 128)[:, 0, :].contiguous().view(B, S, 128)
        # ---- sequence mixer + classifier ----
        for blk in model.sequence_mixer.dilated_convs:
            blk_in = cur
            nl = len(blk.conv_layers)
            for kk, layer in enumerate(blk.conv_layers):
                d = blk.dilations[kk]
                c = K.conv(cur, _f(layer.conv.weight), B, S, S, 128, 128, 7, dil=d, pad=3 * d)
                cur = K.rowln(c, _f(layer.norm.weight.reshape(-1)), _f(layer.norm.bias.reshape(-1)), 1, layer.norm.eps,
                              res=blk_in if kk == nl - 1 else None)
        logits = K.conv(cur, _f(model.classifier.weight), B, S, S, 128, model.num_classes, 1,
                        bias=_f(model.classifier.bias))
    return logits
