"""-m gpu: every CUDA kernel, called through the C ABI, against an fp32 torch restatement of the same op.

Tolerances (written per test): outputs are stored in fp16 (rel. rounding 4.9e-4) and MMA operands are fp16, so the
reference uses fp16-rounded operands with fp32 accumulation; the bound covers fp16 output rounding plus rare
1-ulp operand flips caused by the fast GELU (2.5e-5 abs error) in the prologue.
"""
import ctypes as C

import pytest
import torch

import gpu_utils as G
from wav2sleep_b200 import _lib

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def test_pack_conv_weight_layout(cuda_device):
    w = torch.randn(32, 16, 3, device=cuda_device)
    p = G.pack_conv(w, split=1).view(2, 3, 2, 32, 8)  # [hi|lo][tap][cin/8][cout][8]
    torch.cuda.synchronize()
    ref = w.permute(2, 1, 0).reshape(3, 2, 8, 32).permute(0, 1, 3, 2)
    assert torch.equal(p[0], ref.half())
    assert torch.equal(p[1], (ref - ref.half().float()).half())
    assert (p[0].float() + p[1].float() - ref).abs().max().item() < 1e-6
    # Linear(4C -> F) viewed as taps=4 (input index = tap*C + c)
    wl = torch.randn(128, 4 * 64, device=cuda_device)
    pl = G.pack_conv(wl, taps_major=1, taps=4).view(4, 8, 128, 8)
    ref = wl.view(128, 4, 8, 8).permute(1, 2, 0, 3).half()
    assert torch.equal(pl, ref)


@pytest.mark.parametrize("B,T", [(2, 4096), (3, 1024 + 256), (1, 2048)])
def test_first_conv(cuda_device, B, T):
    """first_conv_kernel vs conv1d(k3,pad1) + 1x1 stride-2 branch + sums + -inf row detection."""
    from wav2sleep_b200 import build_default
    torch.manual_seed(1)
    model = build_default({"ABD": "ABD"}, 4, seed=3).to(cuda_device).eval()
    enc = model.signal_encoders.encoders["ABD"]
    x = torch.randn(B, T, device=cuda_device)
    if B > 1:
        x[1] = float("-inf")
    # run the whole encoder in keep mode and look at block 0's first tensors in the workspace
    from wav2sleep_b200.engine import PackPlan, _PackedEncoder
    lib = _lib.load()
    plan = PackPlan(lib, cuda_device)
    pe = _PackedEncoder(lib, enc, cuda_device, plan)
    plan.run()
    ws_bytes = lib.w2s_encoder_workspace_bytes(C.byref(pe.desc), B, T, 1)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=cuda_device)
    z = torch.zeros(B, T // 256, 128, dtype=torch.float16, device=cuda_device)
    mask = torch.zeros(B, dtype=torch.uint8, device=cuda_device)
    _lib.check(lib.w2s_encoder_fwd(C.byref(pe.desc), x.data_ptr(), B, T, ws.data_ptr(), ws_bytes, 1, z.data_ptr(),
                                   mask.data_ptr(), G.stream()))
    torch.cuda.synchronize()
    assert mask.tolist() == [0] + ([1] if B > 1 else []) + [0] * max(0, B - 2)
    offs = (C.c_int64 * (7 * len(enc.channels)))()
    _lib.check(lib.w2s_encoder_layout(C.byref(pe.desc), B, T, offs))
    s1 = ws[offs[0]: offs[0] + B * 16 * 2 * 8].view(torch.float64).view(B, 16, 2).float()
    y1 = ws[offs[3]: offs[3] + B * T * 16 * 2].view(torch.float16).view(B, T, 16)
    r0 = ws[offs[4]: offs[4] + B * (T // 2) * 16 * 2].view(torch.float16).view(B, T // 2, 16)
    w1 = enc.cnn[0].conv1.conv.weight
    wd = enc.cnn[0].downsample.weight
    live = [b for b in range(B) if not (B > 1 and b == 1)]
    xr = x[live]
    ref = torch.nn.functional.conv1d(xr[:, None], w1, padding=1).transpose(1, 2)
    refr = torch.nn.functional.conv1d(xr[:, None], wd, stride=2).transpose(1, 2)
    assert (y1[live].float() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()  # fp16 storage
    assert (r0[live].float() - refr).abs().max().item() < 2e-3 * refr.abs().max().item()
    assert torch.allclose(s1[live][..., 0], ref.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s1[live][..., 1], (ref * ref).sum(1), rtol=1e-4, atol=1e-2)


ENC_CASES = [
    # cin, cout, stride, has_ds, L
    (16, 16, 1, 0, 2500), (32, 32, 1, 0, 1100), (64, 64, 1, 0, 700), (128, 128, 1, 0, 300),
    (16, 16, 2, 0, 4096), (32, 32, 2, 0, 1030), (64, 64, 2, 0, 512), (128, 128, 2, 0, 258),
    (16, 16, 1, 1, 2048), (16, 32, 1, 1, 1000), (32, 32, 1, 1, 600), (32, 64, 1, 1, 512),
    (64, 64, 1, 1, 256), (64, 128, 1, 1, 384), (128, 128, 1, 1, 130),
]


@pytest.fixture(params=[0, 1], ids=["stream", "tile"])
def conv_impl(request):
    """Run with the persistent streaming kernel (default) and with the tile-per-CTA kernel."""
    lib = _lib.load()
    lib.w2s_set_conv_impl(request.param)
    yield request.param
    lib.w2s_set_conv_impl(0)


@pytest.mark.parametrize("cin,cout,stride,has_ds,L", ENC_CASES)
def test_encoder_conv_layer(cuda_device, conv_impl, cin, cout, stride, has_ds, L):
    """EPI_STATS conv kernels: prologue norm+GELU(+res), k3 conv (tcgen05), fp16 store, sums, fused 1x1."""
    torch.manual_seed(cin * 1000 + cout + stride)
    B = 2
    dev = cuda_device
    y = (torch.randn(B, L, cin, device=dev) * 1.5 + 0.3).half()
    r = torch.randn(B, L, cin, device=dev).half() if has_ds else None
    w = torch.randn(cout, cin, 3, device=dev) / (3 * cin) ** 0.5
    wd = torch.randn(cout, cin, 1, device=dev) / cin ** 0.5 if has_ds else None
    L_out = (L + 2 - 3) // stride + 1
    out = torch.full((B, L_out, cout), float("nan"), dtype=torch.float16, device=dev)
    out_ds = torch.full((B, L_out // 2, cout), float("nan"), dtype=torch.float16, device=dev) if has_ds else None
    stats = torch.zeros(B, cout, 2, device=dev, dtype=torch.float64)
    split = G.uses_split(cin, cout)
    assert split == (1 if max(cin, cout) <= 16 else 0)
    G.run_conv(cin=cin, cout=cout, taps=3, stride=stride, dilation=1, pad=1,
               prologue=_lib.PRO_NORM_RES if has_ds else _lib.PRO_NORM, epilogue=_lib.EPI_STATS, has_ds=has_ds,
               B=B, L_in=L, L_out=L_out, **{"in": y}, in_res=r, in_stats=G.sums(y), w=G.pack_conv(w, split=split),
               w_ds=G.pack_conv(wd, split=split) if has_ds else None, out=out, out_ds=out_ds, out_stats=stats,
               in_eps=1e-2)
    stats = stats.float()
    a = G.prologue_ref(y, r)
    ref = G.conv_ref(a, w, stride=stride, split=split)
    assert ref.shape == out.shape
    assert not torch.isnan(out.float()).any()
    # split layers: only the fp16 rounding of the stored output remains (2^-11 relative to each element)
    assert rel_err(out, ref) < (1e-3 if split else 4e-3)
    assert torch.allclose(stats[..., 0], ref.sum(1), rtol=2e-3, atol=0.05 * L ** 0.5)
    assert torch.allclose(stats[..., 1], (ref * ref).sum(1), rtol=2e-3, atol=1e-2)
    if has_ds:
        refd = G.conv_ref(a, wd, stride=2, pad=0, split=split)
        assert refd.shape == out_ds.shape
        assert rel_err(out_ds, refd) < 4e-3


@pytest.mark.parametrize("cin,cout,stride,has_ds,B,L", [
    (16, 16, 1, 0, 5, 40000), (16, 16, 2, 0, 3, 70000), (16, 32, 1, 1, 4, 33000), (32, 32, 2, 0, 7, 9000),
    (64, 64, 1, 0, 9, 5000), (128, 128, 2, 0, 6, 3000), (128, 128, 1, 1, 5, 2000), (64, 128, 1, 1, 20, 1000),
])
def test_encoder_conv_many_tiles_with_mask(cuda_device, conv_impl, cin, cout, stride, has_ds, B, L):
    """Persistent CTAs walk tile ranges that cross sample boundaries; one sample in the middle is masked out."""
    torch.manual_seed(B * L + cin)
    dev = cuda_device
    y = (torch.randn(B, L, cin, device=dev) + 0.2).half()
    r = torch.randn(B, L, cin, device=dev).half() if has_ds else None
    w = torch.randn(cout, cin, 3, device=dev) / (3 * cin) ** 0.5
    wd = torch.randn(cout, cin, 1, device=dev) / cin ** 0.5 if has_ds else None
    L_out = (L - 1) // stride + 1
    out = torch.zeros(B, L_out, cout, dtype=torch.float16, device=dev)
    out_ds = torch.zeros(B, L_out // 2, cout, dtype=torch.float16, device=dev) if has_ds else None
    stats = torch.zeros(B, cout, 2, device=dev, dtype=torch.float64)
    mask = torch.zeros(B, dtype=torch.uint8, device=dev)
    mask[B // 2] = 1
    split = G.uses_split(cin, cout)
    G.run_conv(cin=cin, cout=cout, taps=3, stride=stride, dilation=1, pad=1,
               prologue=_lib.PRO_NORM_RES if has_ds else _lib.PRO_NORM, epilogue=_lib.EPI_STATS, has_ds=has_ds,
               B=B, L_in=L, L_out=L_out, **{"in": y}, in_res=r, in_stats=G.sums(y), w=G.pack_conv(w, split=split),
               w_ds=G.pack_conv(wd, split=split) if has_ds else None, out=out, out_ds=out_ds, out_stats=stats,
               row_mask=mask, in_eps=1e-2)
    live = [b for b in range(B) if b != B // 2]
    a = G.prologue_ref(y, r)
    ref = G.conv_ref(a, w, stride=stride, split=split)
    assert out[B // 2].abs().max().item() == 0 and stats[B // 2].abs().max().item() == 0
    assert rel_err(out[live], ref[live]) < (1e-3 if split else 4e-3)
    st = stats.float()
    assert torch.allclose(st[live][..., 0], ref[live].sum(1), rtol=2e-3, atol=0.05 * L ** 0.5)
    assert torch.allclose(st[live][..., 1], (ref[live] * ref[live]).sum(1), rtol=2e-3, atol=1e-2)
    if has_ds:
        refd = G.conv_ref(a, wd, stride=2, pad=0, split=split)
        assert out_ds[B // 2].abs().max().item() == 0
        assert rel_err(out_ds[live], refd[live]) < (1e-3 if split else 4e-3)


@pytest.mark.parametrize("cin,cout,stride,has_ds,in_wide,out_wide,block", [
    (16, 16, 1, 0, 1, 1, 1), (16, 16, 2, 0, 1, 1, 1), (16, 32, 1, 1, 1, 0, 2), (16, 32, 1, 1, 1, 1, 2), (32, 32, 1, 0, 1, 1, 2), (32, 32, 2, 0, 1, 1, 3),
    (32, 32, 1, 1, 1, 1, 3), (32, 64, 1, 1, 1, 0, 4), (32, 64, 1, 1, 1, 1, 4), (64, 64, 1, 0, 1, 1, 4), (64, 64, 2, 0, 1, 1, 5),
    (64, 64, 1, 1, 1, 1, 5), (64, 128, 1, 1, 1, 0, 6)])
def test_encoder_conv_wide_storage(cuda_device, cin, cout, stride, has_ds, in_wide, out_wide, block):
    """Wide-storage variants of the streaming kernels (fp32 input / output tensors) incl. the split-operand 64-channel
    kernels of deep encoders (wide_blocks = 6); several tiles per sample, one masked sample."""
    lib = _lib.load()
    torch.manual_seed(cin * 100 + cout + stride + block)
    dev, B, L = cuda_device, 3, 3000
    wide_blocks = (6 if cout == 64 else 4) if out_wide else (2 if cout == 32 else 4)
    split = lib.w2s_encoder_conv_split(wide_blocks, block, cin, cout)
    assert split == (1 if (max(cin, cout) <= 16 or (out_wide and max(cin, cout) <= 64)) else 0)
    y = torch.randn(B, L, cin, device=dev) * 1.5 + 0.3
    r = torch.randn(B, L, cin, device=dev) if has_ds else None
    w = torch.randn(cout, cin, 3, device=dev) / (3 * cin) ** 0.5
    wd = torch.randn(cout, cin, 1, device=dev) / cin ** 0.5 if has_ds else None
    L_out = (L - 1) // stride + 1
    odt = torch.float32 if out_wide else torch.float16
    out = torch.zeros(B, L_out, cout, dtype=odt, device=dev)
    out_ds = torch.zeros(B, L_out // 2, cout, dtype=odt, device=dev) if has_ds else None
    stats = torch.zeros(B, cout, 2, device=dev, dtype=torch.float64)
    mask = torch.tensor([0, 1, 0], dtype=torch.uint8, device=dev)
    yd = y.double()
    in_stats = torch.stack([yd.sum(1), (yd * yd).sum(1)], dim=-1).contiguous()
    G.run_conv(cin=cin, cout=cout, taps=3, stride=stride, dilation=1, pad=1,
               prologue=_lib.PRO_NORM_RES if has_ds else _lib.PRO_NORM, epilogue=_lib.EPI_STATS, has_ds=has_ds,
               B=B, L_in=L, L_out=L_out, **{"in": y}, in_res=r, in_stats=in_stats, w=G.pack_conv(w, split=split),
               w_ds=G.pack_conv(wd, split=split) if has_ds else None, out=out, out_ds=out_ds, out_stats=stats,
               row_mask=mask, in_eps=1e-2, in_wide=in_wide, out_wide=out_wide,
               force_split=1 if (split and not G.uses_split(cin, cout)) else 0)
    mu = y.mean(1, keepdim=True)
    var = (y * y).mean(1, keepdim=True) - mu * mu
    a = G.gelu((y - mu) / torch.sqrt(var.clamp_min(0) + 1e-2))
    if has_ds:
        a = G.gelu(a + r)
    ref = G.conv_ref(a, w, stride=stride, split=split)
    live = [0, 2]
    assert out[1].abs().max().item() == 0 and stats[1].abs().max().item() == 0
    assert rel_err(out[live], ref[live]) < (1e-3 if split else 4e-3)
    st = stats.float()
    assert torch.allclose(st[live][..., 0], ref[live].sum(1), rtol=2e-3, atol=0.05 * L ** 0.5)
    assert torch.allclose(st[live][..., 1], (ref[live] * ref[live]).sum(1), rtol=2e-3, atol=1e-2)
    if has_ds:
        refd = G.conv_ref(a, wd, stride=2, pad=0, split=split)
        assert out_ds[1].abs().max().item() == 0
        assert rel_err(out_ds[live], refd[live]) < (1e-3 if split else 4e-3)


def test_encoder_conv_row_mask_skips_sample(cuda_device):
    dev = cuda_device
    B, L, c = 3, 1024, 16
    y = torch.randn(B, L, c, device=dev).half()
    w = torch.randn(c, c, 3, device=dev) / 7
    out = torch.zeros(B, L, c, dtype=torch.float16, device=dev)
    stats = torch.zeros(B, c, 2, device=dev, dtype=torch.float64)
    mask = torch.tensor([0, 1, 0], dtype=torch.uint8, device=dev)
    G.run_conv(cin=c, cout=c, taps=3, stride=1, dilation=1, pad=1, prologue=_lib.PRO_NORM, epilogue=_lib.EPI_STATS,
               has_ds=0, B=B, L_in=L, L_out=L, **{"in": y}, in_stats=G.sums(y), w=G.pack_conv(w, split=1), out=out,
               out_stats=stats, row_mask=mask, in_eps=1e-2)
    assert out[1].abs().max().item() == 0 and stats[1].abs().max().item() == 0
    assert out[0].abs().max().item() > 0 and out[2].abs().max().item() > 0


@pytest.mark.parametrize("cin,S", [(64, 37), (128, 130), (128, 16)])
def test_encoder_linear(cuda_device, cin, S):
    """Linear(4C -> 128) + bias + GELU as a 4-tap stride-4 conv (reference models/wav2sleep.py:261-265)."""
    torch.manual_seed(cin + S)
    dev, B, L = cuda_device, 2, 4 * S
    y = torch.randn(B, L, cin, device=dev).half()
    r = torch.randn(B, L, cin, device=dev).half()
    w = torch.randn(128, 4 * cin, device=dev) / (4 * cin) ** 0.5
    bias = torch.randn(128, device=dev)
    out = torch.full((B, S, 128), float("nan"), dtype=torch.float16, device=dev)
    G.run_conv(cin=cin, cout=128, taps=4, stride=4, dilation=1, pad=0, prologue=_lib.PRO_NORM_RES,
               epilogue=_lib.EPI_BIAS_GELU, has_ds=0, B=B, L_in=L, L_out=S, **{"in": y}, in_res=r,
               in_stats=G.sums(y), w=G.pack_conv(w, taps_major=1, taps=4), bias=bias, out=out, in_eps=1e-2)
    a = G.prologue_ref(y, r).half().float().reshape(B, S, 4 * cin)
    ref = G.gelu(a @ w.half().float().t() + bias)
    assert rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("dil,S,res", [(1, 200, 0), (4, 130, 0), (32, 300, 0), (32, 1200, 1), (2, 50, 1)])
def test_seq_conv_layer(cuda_device, dil, S, res):
    """Dilated k7 conv + ConvLayerNorm + GELU (+ residual + GELU + classifier head)."""
    torch.manual_seed(dil + S)
    dev, B = cuda_device, 2
    x = torch.randn(B, S, 128, device=dev).half()
    w = torch.randn(128, 128, 7, device=dev) / (7 * 128) ** 0.5
    g, b = torch.randn(128, device=dev) * 0.2 + 1, torch.randn(128, device=dev) * 0.1
    xin = torch.randn(B, S, 128, device=dev).half() if res else None
    hw, hb = torch.randn(5, 128, device=dev) / 11, torch.randn(5, device=dev)
    out = torch.full((B, S, 128), float("nan"), dtype=torch.float16, device=dev)
    logits = torch.full((B, S, 5), float("nan"), device=dev)
    G.run_conv(cin=128, cout=128, taps=7, stride=1, dilation=dil, pad=3 * dil, prologue=_lib.PRO_NONE,
               epilogue=_lib.EPI_LN_GELU_RES if res else _lib.EPI_LN_GELU, has_ds=0, B=B, L_in=S, L_out=S,
               **{"in": x}, w=G.pack_conv(w), ln_w=g, ln_b=b, res=xin, head_w=hw if res else None,
               head_b=hb if res else None, logits=logits if res else None, n_classes=5 if res else 0, out=out,
               ln_eps=1e-5)
    c = G.conv_ref(x, w, stride=1, pad=3 * dil, dil=dil)
    mu = c.mean(-1, keepdim=True)
    var = (c - mu).pow(2).mean(-1, keepdim=True)
    ref = G.gelu((c - mu) / torch.sqrt(var + 1e-5) * g + b)
    if res:
        ref = G.gelu(ref + xin.float())
    assert (out.float() - ref).abs().max().item() < 5e-3
    if res:
        assert (logits - (ref @ hw.t() + hb)).abs().max().item() < 5e-3


@pytest.mark.parametrize("B,S,n_blocks,n_dil,ncls", [(1, 1, 2, 6, 4), (3, 40, 2, 6, 5), (2, 129, 1, 2, 4), (2, 300, 2, 6, 4),
                                                     (3, 1200, 2, 6, 4), (2, 1680, 2, 6, 5), (1, 2048, 2, 3, 4),
                                                     (1, 2100, 2, 6, 4)])
def test_seqmixer_stage(cuda_device, B, S, n_blocks, n_dil, ncls):
    """w2s_seqmixer_head_fwd (SequenceCNN + classifier, models/wav2sleep.py:379-390, 66) against fp32 torch on the
    fp16-rounded operands and inter-layer activations.  S <= 2048 runs the cluster-per-night kernel (seq_mixer.cuh: 1 CTA
    for S <= 128 ... 8 CTAs x 256 rows), S = 2100 the layer-per-launch kernels; both must give the same numbers."""
    lib = _lib.load()
    torch.manual_seed(S + n_dil)
    dev = cuda_device
    x = torch.randn(B, S, 128, device=dev).half()
    d = _lib.SeqDesc()
    d.n_blocks, d.n_dilations, d.kernel_size, d.feature_dim, d.n_classes, d.ln_eps = n_blocks, n_dil, 7, 128, ncls, 1e-5
    keep, ws, gs, bs = [], [], [], []
    for bl in range(n_blocks):
        for k in range(n_dil):
            w = torch.randn(128, 128, 7, device=dev) / (7 * 128) ** 0.5 * 1.5
            g, b = torch.randn(128, device=dev) * 0.2 + 1, torch.randn(128, device=dev) * 0.1
            pw = G.pack_conv(w)
            keep += [pw, g, b]
            ws.append(w), gs.append(g), bs.append(b)
            d.w[bl][k], d.ln_w[bl][k], d.ln_b[bl][k] = pw.data_ptr(), g.data_ptr(), b.data_ptr()
    hw, hb = torch.randn(ncls, 128, device=dev) / 11, torch.randn(ncls, device=dev)
    d.head_w, d.head_b = hw.data_ptr(), hb.data_ptr()
    nbytes = lib.w2s_seqmixer_workspace_bytes(C.byref(d), B, S, 0)
    work = torch.full((nbytes,), 0x7F, dtype=torch.uint8, device=dev)  # garbage (NaN patterns): padding must be written
    feat = torch.full((B, S, 128), float("nan"), dtype=torch.float16, device=dev)
    logits = torch.full((B, S, ncls), float("nan"), device=dev)
    for rep in range(2):  # second call re-uses the dirty workspace
        _lib.check(lib.w2s_seqmixer_head_fwd(C.byref(d), x.data_ptr(), B, S, work.data_ptr(), nbytes, 0, feat.data_ptr(),
                                             logits.data_ptr(), G.stream()))
    torch.cuda.synchronize()
    cur = x.float()
    i = 0
    for bl in range(n_blocks):
        blk_in = cur
        for k in range(n_dil):
            c = G.conv_ref(cur, ws[i], stride=1, pad=3 * 2 ** k, dil=2 ** k)
            mu = c.mean(-1, keepdim=True)
            var = (c - mu).pow(2).mean(-1, keepdim=True)
            a = G.gelu((c - mu) / torch.sqrt(var + 1e-5) * gs[i] + bs[i])
            if k == n_dil - 1:
                a = G.gelu(a + blk_in)
            last = (bl == n_blocks - 1 and k == n_dil - 1)
            final = a
            cur = a if last else a.half().float()
            i += 1
    ref_logits = final @ hw.t() + hb
    assert torch.isfinite(logits).all() and torch.isfinite(feat.float()).all()
    e_feat = (feat.float() - final).abs().max().item()
    e_log = (logits - ref_logits).abs().max().item()
    print(f"seq mixer B={B} S={S}: feature max-abs {e_feat:.3e}, logits max-abs {e_log:.3e}")
    # fp16 re-rounding of 1-ulp-different intermediates over n_blocks * n_dil layers
    assert e_feat < 1.5e-2 and e_log < 1.5e-2


@pytest.mark.parametrize("c", [32, 64, 128])
def test_bf16_operands_on_hardware(cuda_device, c):
    """The north star names bf16; this path multiplies fp16 operands.  Hardware evidence at the kernel level: the same
    encoder conv (InstanceNorm + GELU prologue, k3, fp32 accumulate in TMEM) with the MMA operand type flipped to bf16
    (w2s_set_conv_impl(3): bf16 rounding in the prologue, bf16-packed weights, bf16 instruction descriptor), both
    against fp32 torch on UNROUNDED operands.  bf16 keeps 8 mantissa bits instead of 11: its operand error is ~8x the
    fp16 one, on top of the common fp16 output rounding (the model-level consequence - 7.6e-2 / 99.4 % against
    1.05e-2 / 100 % on two cardio nights - is tools/emulate_16bit.py, profiles/r02_emulation_16bit.txt)."""
    lib = _lib.load()
    torch.manual_seed(c)
    dev, B, L = cuda_device, 2, 4096
    y = torch.randn(B, L, c, device=dev).half()
    w = torch.randn(c, c, 3, device=dev) / (3 * c) ** 0.5
    ref = G.conv_ref(G.prologue_ref(y), w, split=1)  # fp32, operands not rounded
    err = {}
    try:
        for name, impl, split in (("fp16", 1, 0), ("bf16", 3, 2)):
            _lib.check(lib.w2s_set_conv_impl(impl))
            out = torch.full((B, L, c), float("nan"), dtype=torch.float16, device=dev)
            stats = torch.zeros(B, c, 2, dtype=torch.float64, device=dev)
            G.run_conv(cin=c, cout=c, taps=3, stride=1, dilation=1, pad=1, prologue=_lib.PRO_NORM, epilogue=_lib.EPI_STATS,
                       has_ds=0, B=B, L_in=L, L_out=L, **{"in": y}, in_stats=G.sums(y), w=G.pack_conv(w, split=split),
                       out=out, out_stats=stats, in_eps=1e-2)
            assert torch.isfinite(out.float()).all()
            err[name] = ((out.float() - ref).abs().mean() / ref.abs().mean()).item()
    finally:
        _lib.check(lib.w2s_set_conv_impl(0))
    print(f"conv {c}->{c}: mean relative error with fp16 operands {err['fp16']:.3e}, with bf16 operands {err['bf16']:.3e} "
          f"({err['bf16'] / err['fp16']:.1f}x)")
    assert err["fp16"] < 1e-3 and err["bf16"] < 1e-2   # both compute the same convolution ...
    assert err["bf16"] > 2.5 * err["fp16"]              # ... bf16 operands are markedly less accurate


def test_pack_batch_matches_single_packs(cuda_device):
    """w2s_pack_batch (strided sources, one launch) against w2s_pack_conv_weight / w2s_pack_linear_frag, incl. the
    flipped + transposed view used for data-gradient weights and the taps-major view of a Linear."""
    from wav2sleep_b200.engine import PackPlan
    lib = _lib.load()
    torch.manual_seed(3)
    w = torch.randn(32, 16, 3, device=cuda_device)
    lin = torch.randn(128, 256, device=cuda_device)  # Linear(4 * 64 -> 128)
    fr = torch.randn(384, 128, device=cuda_device)
    plan = PackPlan(lib, cuda_device)
    jobs = [(plan.conv(w, split=1), w, 0, 1), (plan.conv(w.permute(1, 0, 2), flip=True), w.flip(2).permute(1, 0, 2).contiguous(), 0, 0),
            (plan.conv(lin.view(128, 4, 64).permute(0, 2, 1)), lin, 1, 0)]
    p_frag = plan.frag(fr)
    plan.run()
    torch.cuda.synchronize()
    bufs = {t.data_ptr(): t for t in plan.keep}
    for ptr, src, taps_major, split in jobs:
        if taps_major:
            cout, cin, taps = src.shape[0], src.shape[1] // 4, 4
        else:
            cout, cin, taps = src.shape
        ref = torch.empty(cout * cin * taps * (2 if split else 1), dtype=torch.float16, device=cuda_device)
        _lib.check(lib.w2s_pack_conv_weight(src.data_ptr(), cout, cin, taps, taps_major, split, ref.data_ptr(), G.stream()))
        torch.cuda.synchronize()
        assert torch.equal(bufs[ptr], ref)
    ref = torch.empty(fr.numel(), dtype=torch.float16, device=cuda_device)
    _lib.check(lib.w2s_pack_linear_frag(fr.data_ptr(), 384, 128, ref.data_ptr(), G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(bufs[p_frag], ref)
    # values change in place -> run() again picks them up without rebuilding
    w.mul_(2.0)
    plan.run()
    ref = torch.empty(32 * 16 * 3 * 2, dtype=torch.float16, device=cuda_device)
    _lib.check(lib.w2s_pack_conv_weight(w.data_ptr(), 32, 16, 3, 0, 1, ref.data_ptr(), G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(bufs[jobs[0][0]], ref)
