"""CPU emulation of the CUDA path's 16-bit decisions inside the oracle graph (evidence behind DESIGN.md "Numerics").

Every place where the CUDA forward rounds is rounded here the same way, in the fp32 CPU graph of the oracle:
  * conv OUTPUT storage (pre-norm y, residual branch r): fp16 / bf16 / fp32 per encoder block ("wide" blocks = fp32),
  * conv OPERANDS (activated input a, weights w): one 16-bit value (fp16 / bf16) or a hi + lo pair ("split", ~22 bits),
  * encoder Linear, epoch mixer and sequence mixer operands / stored activations: 16-bit or fp32,
  * GELU: exact erf or the fitted tanh form of the conv prologues.
The statistics of InstanceNorm are taken from the un-rounded accumulators (as the kernels do) and applied to the stored
values.  Usage (prints max-abs / mean logit error and argmax agreement against the plain fp32 oracle):

    python tools/emulate_16bit.py cardio            # fp16 vs bf16 decision, cardio model, 2 nights
    python tools/emulate_16bit.py eog [nights]      # EOG model (14-h nights): which layers need more than fp16
    python tools/emulate_16bit.py grad              # gradient error caused by a 16-bit FORWARD alone (fp32 backward)
"""
import math
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import make_inputs  # noqa: E402
from oracle import wav2sleep_oracle as oracle  # noqa: E402
from wav2sleep_b200 import build_default  # noqa: E402


def rnd(x, kind):
    if kind == "f32" or kind == "split":  # a hi + lo fp16 pair carries ~22 bits: treated as exact
        return x
    if kind == "f16":
        return x.half().float()
    if kind == "bf16":
        return x.bfloat16().float()
    raise ValueError(kind)


def gelu_tanh_fit(x, approx_amp=0.0):
    """The conv prologues' GELU: 0.5 x (1 + tanh(x q(min(x^2, 25)))).  approx_amp > 0 models MUFU.TANH (tanh.approx.f32,
    max relative error 2^-11): a deterministic, oscillating relative error of that amplitude on the tanh value."""
    t = (x * x).clamp(max=25.0)
    q = (t * -3.5159264e-4 + 0.037005995) * t + 0.79750759
    u = x * q
    th = torch.tanh(u)
    if approx_amp > 0:
        th = th * (1.0 + approx_amp * torch.sin(u * 4096.0))
    return 0.5 * x * (1.0 + th)


def gelu_half(x, tanh_ulp=0.0):
    """The prologue GELU evaluated in fp16 arithmetic (HMUL2 / HFMA2 / tanh.approx.f16x2: every operation rounds to
    fp16, x_hat itself is rounded first).  tanh_ulp > 0 adds a deterministic error of that many fp16 ulps (of 1.0) to
    the tanh value, modelling tanh.approx.f16x2's stated max absolute error of 2^-10.987."""
    h = lambda v: v.half().float()
    c2, c1, c0 = (h(torch.tensor(c)) for c in (-3.5159264e-4, 0.037005995, 0.79750759))
    x = h(x)
    t = h(x * x).clamp(max=25.0)
    q = h(t * c2 + c1)
    q = h(q * t + c0)
    u = h(x * q)
    th = torch.tanh(u)
    if tanh_ulp > 0:
        th = th + tanh_ulp * 2.0 ** -11 * torch.sin(u * 4096.0)
    th = h(th)
    hx = 0.5 * x
    return h(hx * th + hx)


class Policy:
    def __init__(self, store="f16", wide_blocks=0, op="f16", split_max=32, lin="f16", mixer="f16", seq="f16",
                 gelu="tanh", z_store="f16", split_kind="split", split_min=0, half_min=0, tanh_ulp=0.0):
        self.half_min, self.tanh_ulp = half_min, tanh_ulp  # half_min > 0: prologues of kernels with CIN >= half_min use fp16 math
        self.store, self.wide_blocks, self.op, self.split_max = store, wide_blocks, op, split_max
        self.split_kind, self.split_min = split_kind, split_min
        self.lin, self.mixer, self.seq, self.gelu, self.z_store = lin, mixer, seq, gelu, z_store

    def store_kind(self, block):
        return "f32" if block < self.wide_blocks else self.store

    def op_kind(self, cin, cout):
        lo, hi = min(cin, cout), max(cin, cout)
        return self.split_kind if (hi <= self.split_max and lo >= self.split_min) else self.op

    def act(self, x, cin=0):
        if self.half_min and cin >= self.half_min:
            return gelu_half(x, self.tanh_ulp)
        if self.gelu == "tanh_std":  # textbook tanh form (no clamp needed: monotone argument), max abs error ~3e-4 vs erf
            return 0.5 * x * (1.0 + torch.tanh(0.7978845608 * (x + 0.044715 * x * x * x)))
        if self.gelu == "tanh_approx":
            return gelu_tanh_fit(x, approx_amp=2.0 ** -11.5)
        return gelu_tanh_fit(x) if self.gelu == "tanh" else oracle.gelu(x)

    def __repr__(self):
        return (f"store={self.store} wide={self.wide_blocks} op={self.op} {self.split_kind}<={self.split_max} lin={self.lin} "
                f"mixer={self.mixer} seq={self.seq} gelu={self.gelu}"
                + (f" half-gelu>={self.half_min} tanh_ulp={self.tanh_ulp}" if self.half_min else ""))


def norm_apply(y_acc, y_st, eps):
    """InstanceNorm with statistics of the accumulators applied to the stored values."""
    mu = y_acc.double().mean(2, keepdim=True)
    var = (y_acc.double() ** 2).mean(2, keepdim=True) - mu * mu
    return ((y_st.double() - mu) / torch.sqrt(var.clamp_min(0) + eps)).float()


def conv(a, w, kind, stride=1, pad=1):
    """kind 'wsplit' / 'asplit': only the weights / only the activations are carried as hi + lo pairs (2 MMAs, not 3)."""
    ka = "f16" if kind == "wsplit" else ("f32" if kind == "asplit" else kind)
    kw = "f16" if kind == "asplit" else ("f32" if kind == "wsplit" else kind)
    return F.conv1d(rnd(a, ka), rnd(w, kw), None, stride=stride, padding=pad)


def encoder(x_BT, sd, prefix, spe, eps, pol: Policy):
    B = x_BT.size(0)
    nb = int(math.log2(spe)) - 2
    a_in = x_BT.unsqueeze(1)  # block input (activated), fp32 for block 0
    cin = 1
    for i in range(nb):
        p = f"{prefix}.cnn.{i}"
        w1, w2, w3, wd = (sd[f"{p}.conv1.conv.weight"], sd[f"{p}.conv2.conv.weight"], sd[f"{p}.conv3.conv.weight"],
                          sd[f"{p}.downsample.weight"])
        c = w1.shape[0]
        st = pol.store_kind(i)
        k1 = "f32" if i == 0 else pol.op_kind(cin, c)  # block 0 conv1 / downsample run in fp32 from the raw signal
        y1 = conv(a_in, w1, k1)
        r = rnd(conv(a_in, wd, k1, stride=2, pad=0), "f32" if i == 0 else st)  # block 0: recomputed, never stored
        y1s = y1 if i == 0 else rnd(y1, st)       # block-0 conv1 is recomputed in the consumer: never stored
        a1 = pol.act(norm_apply(y1, y1s, eps), c)
        k = pol.op_kind(c, c)
        y2 = conv(a1, w2, k)
        a2 = pol.act(norm_apply(y2, rnd(y2, st), eps), c)
        y3 = conv(a2, w3, k, stride=2)
        last = i == nb - 1  # the last block's output is activated in the Linear's prologue (fp32 math)
        a3 = pol.act(norm_apply(y3, rnd(y3, st), eps), 0 if last else c)
        a_in = pol.act(a3 + r, 0 if last else c)
        cin = c
    C = a_in.size(1)
    y = rnd(a_in, pol.lin).transpose(1, 2).reshape(B, -1, 4 * C)
    z = y @ rnd(sd[f"{prefix}.linear.weight"], pol.lin).t() + sd[f"{prefix}.linear.bias"]
    return rnd(oracle.gelu(z), pol.z_store)


def mixer_layer(x, key_mask, sd, prefix, nhead, eps, kind):
    N, D, Fd = x.shape
    hd = Fd // nhead
    W = lambda n: rnd(sd[f"{prefix}.{n}"], kind)
    h = rnd(oracle.layer_norm(x, sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], eps), kind)
    qkv = rnd(h @ W("self_attn.in_proj_weight").t() + sd[f"{prefix}.self_attn.in_proj_bias"], kind)
    q, k, v = qkv.split(Fd, dim=-1)
    q, k, v = (t.view(N, D, nhead, hd).transpose(1, 2) for t in (q, k, v))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    s = s.masked_fill(key_mask[:, None, None, :], float("-inf"))
    a = rnd((torch.softmax(s, -1) @ v).transpose(1, 2).reshape(N, D, Fd), kind)
    x = x + a @ W("self_attn.out_proj.weight").t() + sd[f"{prefix}.self_attn.out_proj.bias"]
    h = rnd(oracle.layer_norm(x, sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"], eps), kind)
    h = rnd(oracle.gelu(h @ W("linear1.weight").t() + sd[f"{prefix}.linear1.bias"]), kind)
    return x + h @ W("linear2.weight").t() + sd[f"{prefix}.linear2.bias"]


def forward(x, sd, cfg, pol: Policy):
    z = {}
    for name, x_BT in x.items():
        enc = cfg.signal_map[name]
        z[name] = encoder(x_BT, sd, f"signal_encoders.encoders.{enc}", oracle.SAMPLES_PER_EPOCH[name],
                          cfg.instance_eps, pol)
    names = sorted(z)
    B, S, Fd = z[names[0]].shape
    cls = sd["epoch_mixer.register_tokens"][0, 0, :, 0]
    t = torch.stack([cls.expand(B, S, Fd)] + [z[n] for n in names], dim=2).reshape(B * S, len(names) + 1, Fd)
    km = torch.zeros(B * S, len(names) + 1, dtype=torch.bool)
    for l in range(cfg.mixer_layers):
        t = mixer_layer(t, km, sd, f"epoch_mixer.transformer_encoder.layers.{l}", cfg.nhead, cfg.layer_eps, pol.mixer)
    m = rnd(t[:, 0, :].reshape(B, S, Fd), pol.mixer)
    xs = m.transpose(1, 2)
    for bl in range(cfg.seq_blocks):
        out = xs
        for k in range(cfg.seq_dilations):
            pre = f"sequence_mixer.dilated_convs.{bl}.conv_layers.{k}"
            d = 2 ** k
            out = F.conv1d(rnd(out, pol.seq), rnd(sd[f"{pre}.conv.weight"], pol.seq), None, padding=3 * d, dilation=d)
            out = oracle.gelu(oracle.channel_layer_norm(out, sd[f"{pre}.norm.weight"], sd[f"{pre}.norm.bias"], cfg.layer_eps))
            if k < cfg.seq_dilations - 1:
                out = rnd(out, pol.seq)
        xs = rnd(oracle.gelu(out + xs), pol.seq)
    return xs.transpose(1, 2) @ sd["classifier.weight"].t() + sd["classifier.bias"]


def report(tag, out, ref):
    err = (out - ref).abs()
    agree = (out.argmax(-1) == ref.argmax(-1)).float().mean().item()
    flips = int((out.argmax(-1) != ref.argmax(-1)).sum())
    print(f"{tag:95s} max {err.max().item():.3e} mean {err.mean().item():.3e} argmax {100 * agree:.3f}% ({flips} flips)",
          flush=True)


def grad_floor():
    """Parameter-gradient error of an fp32 backward through a forward whose activations / operands are rounded like the
    CUDA path's (straight-through rounding): the floor of any 16-bit-forward training path.  One 10-h ECG night, the
    loss divided by 16 as inside a batch of 16 nights."""
    global rnd
    exact = rnd

    def rnd_st(x, kind):
        r = exact(x, kind)
        return x if r is x else x + (r - x).detach()

    rnd = rnd_st
    sig, S = {"ECG": "ECG"}, 1200
    torch.manual_seed(0)
    model = build_default(sig, 4, seed=0)
    x = make_inputs(sig, 1, S, seed=11)
    labels = torch.randint(0, 4, (1, S))
    cfg = oracle.OracleConfig(signal_map=sig, num_classes=4)

    def grads(fwd):
        params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        loss = F.cross_entropy(fwd(params).view(-1, 4), labels.view(-1)) / 16
        loss.backward()
        return {k: v.grad for k, v in params.items()}

    g_ref = grads(lambda p: oracle.forward_with_grad(x, p, cfg))
    for pol in (Policy(split_max=16), Policy(split_max=32)):
        g = grads(lambda p: forward(x, p, cfg, pol))
        rows = sorted(((g[k] - g_ref[k]).norm().item() / max(g_ref[k].norm().item(), 1e-20), k) for k in g_ref)
        print(pol)
        for rel, k in rows[-5:][::-1]:
            print(f"    {k:62s} rel {rel:.3e}")
        print(f"    median over {len(rows)} tensors {rows[len(rows) // 2][0]:.3e}", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "grad":
        return grad_floor()
    with torch.no_grad():
        return _main()


def _main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cardio"
    if which == "cardio":
        smap, ncls, S, nights, cfg = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}, 4, 1200, 2, oracle.cardio_config()
        pols = [Policy(store="f16", op="f16", split_max=0), Policy(store="bf16", op="bf16", split_max=0, lin="bf16",
                                                                   mixer="bf16", seq="bf16", z_store="bf16"),
                Policy(store="f16", op="f16", split_max=32),
                Policy(store="f16", op="f16", split_max=32, split_kind="wsplit"),
                Policy(store="f16", op="f16", split_max=32, split_kind="asplit")]
        if len(sys.argv) > 3:
            nights = int(sys.argv[2])
            pols = [eval(sys.argv[3])]
    else:
        smap, ncls, S, cfg = {"EOG-L": "EOG-L", "EOG-R": "EOG-R"}, 5, 1680, oracle.eog_config()
        nights = int(sys.argv[2]) if len(sys.argv) > 2 else 1
        pols = [
            Policy(wide_blocks=4),                                              # round-1 default
            Policy(wide_blocks=4, mixer="f32", seq="f32", lin="f32", z_store="f32"),   # what the mixers/linear cost
            Policy(wide_blocks=6, split_max=64),                                # + 64-channel blocks wide + split
            Policy(wide_blocks=10, split_max=64),                               # + fp32 storage of the 128-channel blocks
            Policy(wide_blocks=10, split_max=128),                              # + split 128-channel operands
            Policy(wide_blocks=10, split_max=128, gelu="erf"),                  # + exact GELU
            Policy(wide_blocks=10, split_max=128, gelu="erf", mixer="f32", seq="f32", lin="f32", z_store="f32"),
        ]
        if len(sys.argv) > 3:
            pols = [eval(a) for a in sys.argv[3:]]  # any number of Policy(...) expressions
    model = build_default(smap, ncls, seed=0)
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    x = make_inputs(smap, nights, S, seed=42)
    t0 = time.time()
    ref = oracle.forward(x, sd, cfg)
    print(f"{which}: {nights} night(s), oracle {time.time() - t0:.1f} s", flush=True)
    for pol in pols:
        report(repr(pol), forward(x, sd, cfg, pol), ref)


if __name__ == "__main__":
    main()
