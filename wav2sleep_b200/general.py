"""General inference engine: every model option of the reference that is outside the fused tcgen05 path
 on the dimension-generic fp32 kernels of ``csrc/general.cuh``.

What is restated, with the reference lines it follows (relative to /root/reference/src/wav2sleep/):
  SignalEncoders.forward          models/wav2sleep.py:146-161   (-inf rows, optional signal embedding :155-159)
  SignalEncoder.forward           models/wav2sleep.py:235-267   (chunk-causal per-epoch path :248-255, output_norm :266)
  ConvBlock1D / ConvLayer1D       models/blocks.py:57-71, 173-186 (causal padding + right-trim :150-152, :178-182)
  norms / activations             models/utils.py:9-96
  MultiModalAttentionEmbedder     models/wav2sleep.py:301-346   (register tokens :299, :330-335; norm_first either way)
  SequenceCNN / DilatedConvBlock  models/wav2sleep.py:379-390 ; models/blocks.py:115-126

Tensors are fp32, channels-last ``[B, L, C]``; weights stay in their PyTorch layouts (no packing).  torch is used for
allocation and pure data movement (views, stack, index) only - every FLOP is in ``libw2s_b200.so``.  Inference only:
BatchNorm uses its running statistics, dropout is the identity (``model.eval()`` semantics); the training path exists
for the default model family (training.py).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import _lib
from .engine import _stream

ACT = {"linear": 0, "relu": 1, "leaky": 2, "gelu": 3, "silu": 4, "swish": 4}
NORM_INSTANCE, NORM_GROUP, NORM_BATCH = 0, 1, 2


def _f(t: Tensor | None):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _p(t: Tensor | None):
    return None if t is None else t.data_ptr()


class GeneralEngine:
    def __init__(self, model):
        self.model = model
        self.lib = _lib.load()

    # ------------------------------------------------------------------ kernel wrappers
    def conv(self, x: Tensor, w: Tensor, bias, L_out: int, stride=1, dil=1, pad_left=0, mask=None, taps_major=0, raw=0):
        """x [B, L_in, cin] -> [B, L_out, cout]; w [cout, cin, taps] (taps_major: [cout, taps * cin] with taps given by
        the caller through w.shape[1] // cin)."""
        B, L_in, cin = x.shape
        w = _f(w)
        if taps_major:
            cout, taps = w.shape[0], w.shape[1] // cin
        else:
            cout, _, taps = w.shape
        out = torch.zeros(B, L_out, cout, dtype=torch.float32, device=x.device)
        _lib.check(self.lib.w2s_gen_conv(x.data_ptr(), w.data_ptr(), _p(_f(bias)), out.data_ptr(), _p(mask), B, L_in, L_out,
                                         cin, cout, taps, stride, dil, pad_left, taps_major, raw, _stream()))
        return out

    def linear(self, x: Tensor, lin: nn.Linear | None = None, weight=None, bias=None) -> Tensor:
        """rows [R, K] -> [R, N] = x W^T + b (a 1-tap conv over R positions)."""
        if lin is not None:
            weight, bias = lin.weight, lin.bias
        R, K = x.shape
        return self.conv(x.view(1, R, K), _f(weight).unsqueeze(-1), bias, R).view(R, -1)

    def affine_act(self, x: Tensor, act: str, scale=None, shift=None, res=None, mask=None, per_channel=0) -> Tensor:
        B, L, Cc = x.shape
        out = torch.zeros_like(x)
        _lib.check(self.lib.w2s_gen_affine_act(x.data_ptr(), _p(scale), _p(shift), _p(res), out.data_ptr(), _p(mask), B, L,
                                               Cc, ACT[act], per_channel, _stream()))
        return out

    def rownorm(self, x: Tensor, weight, bias, eps: float, act: str = "linear", rms: int = 0) -> Tensor:
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        out = torch.empty_like(x)
        _lib.check(self.lib.w2s_gen_rownorm(x.data_ptr(), _p(_f(weight)), _p(_f(bias)), out.data_ptr(), rows, Cc, rms,
                                            ACT[act], eps, _stream()))
        return out

    def norm_act(self, y: Tensor, layer, act: str, mask=None, res=None) -> Tensor:
        """norm (by the layer's kind) -> activation of a conv output y [B, L, C]   (ConvLayer1D.forward, blocks.py:183-185)."""
        kind, norm = layer.norm_name, layer.norm
        B, L, Cc = y.shape
        if kind is None:
            return self.affine_act(y, act, mask=mask)
        if kind in ("layer", "rms"):
            w = norm.weight.reshape(-1)
            b = norm.bias.reshape(-1) if kind == "layer" else None
            return self.rownorm(y, w, b, norm.eps, act, rms=int(kind == "rms"))
        dev = y.device
        if kind == "batch":  # eval mode: running statistics
            scale = torch.empty(1, Cc, dtype=torch.float32, device=dev)
            shift = torch.empty(1, Cc, dtype=torch.float32, device=dev)
            _lib.check(self.lib.w2s_gen_norm_consts(None, _p(_f(norm.weight)), _p(_f(norm.bias)), _f(norm.running_mean).data_ptr(),
                                                    _f(norm.running_var).data_ptr(), scale.data_ptr(), shift.data_ptr(), 1, Cc, L,
                                                    NORM_BATCH, 1, norm.eps, _stream()))
            return self.affine_act(y, act, scale, shift, mask=mask, per_channel=1)
        stats = torch.zeros(B, Cc, 2, dtype=torch.float64, device=dev)
        _lib.check(self.lib.w2s_gen_stats(y.data_ptr(), stats.data_ptr(), _p(mask), B, L, Cc, _stream()))
        scale = torch.zeros(B, Cc, dtype=torch.float32, device=dev)
        shift = torch.zeros(B, Cc, dtype=torch.float32, device=dev)
        if kind == "instance":
            w = _f(norm.weight) if getattr(norm, "affine", False) else None
            b = _f(norm.bias) if getattr(norm, "affine", False) else None
            _lib.check(self.lib.w2s_gen_norm_consts(stats.data_ptr(), _p(w), _p(b), None, None, scale.data_ptr(),
                                                    shift.data_ptr(), B, Cc, L, NORM_INSTANCE, 1, norm.eps, _stream()))
        elif kind == "group":
            gn = norm.norm
            _lib.check(self.lib.w2s_gen_norm_consts(stats.data_ptr(), _p(_f(gn.weight)), _p(_f(gn.bias)), None, None,
                                                    scale.data_ptr(), shift.data_ptr(), B, Cc, L, NORM_GROUP, gn.num_groups,
                                                    gn.eps, _stream()))
        else:
            raise NotImplementedError(f"norm {kind!r}")
        return self.affine_act(y, act, scale, shift, mask=mask)

    def conv_layer(self, a: Tensor, layer, mask=None, raw=0) -> Tensor:
        """ConvLayer1D.forward (blocks.py:173-186) on a [B, L, Cin] input."""
        L_out = layer.output_length(a.shape[1])
        y = self.conv(a, layer.conv.weight, layer.conv.bias, L_out, stride=layer.stride, dil=layer.dilation,
                      pad_left=layer.padding, mask=mask, raw=raw)
        return self.norm_act(y, layer, layer.activation_name, mask=mask)

    def conv_block(self, a: Tensor, blk, mask=None, raw=0) -> Tensor:
        """ConvBlock1D.forward (blocks.py:57-71): act(conv3(conv2(conv1(x))) [+ downsample(x)])."""
        y = self.conv_layer(a, blk.conv1, mask, raw)
        y = self.conv_layer(y, blk.conv2, mask)
        y = self.conv_layer(y, blk.conv3, mask)
        if not blk.use_residual:
            return self.affine_act(y, blk.activation_name, mask=mask)
        r = self.conv(a, blk.downsample.weight, None, (a.shape[1] - 1) // 2 + 1, stride=2, mask=mask, raw=raw)
        if r.shape[1] != y.shape[1]:
            raise ValueError(f"residual length {r.shape[1]} != conv path length {y.shape[1]}")
        return self.affine_act(y, blk.activation_name, res=r, mask=mask)

    def dilated_blocks(self, x_BSF: Tensor, blocks) -> Tensor:
        """Sequential(DilatedConvBlock) in eval mode (blocks.py:115-126)."""
        cur = x_BSF
        for blk in blocks:
            blk_in = cur
            for layer in blk.conv_layers:
                cur = self.conv_layer(cur, layer)
            cur = self.affine_act(cur, blk.activation_name, res=blk_in)
        return cur

    # ------------------------------------------------------------------ stages
    def encode(self, enc, x_BT: Tensor):
        """SignalEncoder.forward (wav2sleep.py:235-267) -> (z [B, S, F], mask [B] uint8)."""
        B, T = x_BT.shape
        spe = enc.samples_per_epoch
        if T % spe:
            raise ValueError(f"Input length {T} must be divisible by self.samples_per_epoch={spe}.")
        S = T // spe
        xs = _f(x_BT)
        mask = torch.isinf(xs[:, 0]).to(torch.uint8).contiguous()  # data-dependent control, no arithmetic
        if enc.causal and enc.chunk_causal:  # every epoch is its own sample (wav2sleep.py:248-255)
            a = xs.view(B * S, spe, 1)
            m = mask.repeat_interleave(S).contiguous()
        else:
            a = xs.view(B, T, 1)
            m = mask
        raw = 1
        for blk in enc.cnn:
            a = self.conv_block(a, blk, m, raw)
            raw = 0
        Bp, L4, Cc = a.shape
        if L4 % 4 or (Bp * L4) // 4 != B * S:
            raise ValueError(f"encoder output length {L4} does not give 4 frames per epoch")
        z = self.conv(a, enc.linear.weight, enc.linear.bias, L4 // 4, stride=4, mask=m, taps_major=1)  # [Bp, L4/4, F]
        z = self.affine_act(z, enc.activation_name, mask=m)
        if isinstance(enc.output_norm, nn.LayerNorm):
            z = self.rownorm(z, enc.output_norm.weight, enc.output_norm.bias, enc.output_norm.eps)
        return z.view(B, S, -1), mask

    def epoch_mixer(self, zs: list[Tensor], masks: list[Tensor], B: int, S: int) -> Tensor:
        """MultiModalAttentionEmbedder.forward (wav2sleep.py:301-346) -> CLS features [B, S, F]."""
        mix = self.model.epoch_mixer
        Fd, R = mix.feature_dim, mix.num_register_tokens
        N, dev = B * S, zs[0].device
        if zs[0].shape[-1] != Fd:
            raise ValueError(f"Feature dimension {zs[0].shape[-1]} does not match self.feature_dim={Fd}.")
        reg = _f(mix.register_tokens[0, 0]).t().contiguous()  # [R + 1, F]; token 0 is the CLS token
        toks = [reg[i].expand(B, S, Fd) for i in range(R + 1)]
        toks += [torch.where(m.bool()[:, None, None], torch.zeros_like(z), z) for z, m in zip(zs, masks)]
        D = len(toks)
        x = torch.stack(toks, dim=2).reshape(N * D, Fd).contiguous()
        km = torch.stack([torch.zeros_like(masks[0])] * (R + 1) + masks, dim=1)[:, None, :].expand(B, S, D)
        km = km.reshape(N * D).contiguous()
        H = mix.nhead
        act = mix.activation_name
        for layer in mix.transformer_encoder.layers:
            def sa(h):
                Win, bin_ = _f(layer.self_attn.in_proj_weight), _f(layer.self_attn.in_proj_bias)
                q, k, v = (self.linear(h, weight=Win[j * Fd:(j + 1) * Fd].contiguous(), bias=bin_[j * Fd:(j + 1) * Fd].contiguous())
                           for j in range(3))
                o = torch.empty_like(q)
                _lib.check(self.lib.w2s_gen_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), km.data_ptr(), N, D, H,
                                                 Fd // H, _stream()))
                return self.linear(o, layer.self_attn.out_proj)

            def ff(h):
                hid = self.affine_act(self.linear(h, layer.linear1).unsqueeze(0), act).squeeze(0)
                return self.linear(hid, layer.linear2)

            add = lambda a, b: self.affine_act(a.unsqueeze(0), "linear", res=b.unsqueeze(0)).squeeze(0)
            ln = lambda t, nrm: self.rownorm(t, nrm.weight, nrm.bias, nrm.eps)
            if mix.norm_first:
                x = add(sa(ln(x, layer.norm1)), x)
                x = add(ff(ln(x, layer.norm2)), x)
            else:
                x = ln(add(sa(x), x), layer.norm1)
                x = ln(add(ff(x), x), layer.norm2)
        return x.view(N, D, Fd)[:, 0, :].contiguous().view(B, S, Fd)

    def sequence_mixer(self, x_BSF: Tensor) -> Tensor:
        """SequenceCNN.forward (wav2sleep.py:379-390) + DilatedConvBlock.forward (blocks.py:115-126), eval mode."""
        return self.dilated_blocks(x_BSF, self.model.sequence_mixer.dilated_convs)

    # ------------------------------------------------------------------ model
    @torch.no_grad()
    def forward(self, x: dict[str, Tensor]) -> Tensor:
        m = self.model
        if not isinstance(x, dict) or len(x) == 0:
            raise ValueError("No signals provided to MultiModalAttentionEmbedder.")
        names = sorted(x.keys())  # token order of the mixer, wav2sleep.py:311
        zs, masks = [], []
        B = S = None
        for n in names:
            if n not in m.signal_encoders.signal_map:
                raise KeyError(f"Signal {n!r} has no encoder (valid: {list(m.signal_encoders.signal_map)})")
            t = x[n]
            if not isinstance(t, Tensor) or t.dim() != 2:
                raise ValueError(f"{n}: expected a [B, T] tensor")
            if not t.is_cuda:
                raise RuntimeError("wav2sleep_b200 runs on CUDA (sm_100a) only: move inputs with .to('cuda'); "
                                   "there is no CPU fallback")
        dev = x[names[0]].device
        with torch.cuda.device(dev):
            for n in names:
                t = x[n]
                z, mask = self.encode(m.signal_encoders.get_encoder(n), t)
                if B is None:
                    B, S = z.shape[0], z.shape[1]
                elif (z.shape[0], z.shape[1]) != (B, S):
                    raise ValueError(f"{n}: batch/epoch count {tuple(z.shape[:2])} differs from {(B, S)}")
                if m.signal_encoders.embed_signals:  # wav2sleep.py:155-159
                    e = _f(m.signal_encoders.embedder.weight[m.signal_encoders.sig_to_embedding_idx[n]])
                    z = self.affine_act(z, "linear", shift=e, mask=mask, per_channel=1)
                zs.append(z)
                masks.append(mask)
            feat = self.sequence_mixer(self.epoch_mixer(zs, masks, B, S))
            logits = self.linear(feat.view(B * S, -1), m.classifier).view(B, S, m.num_classes)
        return logits

    @torch.no_grad()
    def predict(self, x: dict[str, Tensor]) -> Tensor:
        logits = self.forward(x)
        B, S, Cn = logits.shape
        out = torch.empty(B, S, dtype=torch.int64, device=logits.device)
        with torch.cuda.device(logits.device):
            _lib.check(self.lib.w2s_argmax(logits.data_ptr(), B * S, Cn, out.data_ptr(), _stream()))
        return out
