"""Generate tests/golden/train_dropout_cardio.npz: the REAL reference model in train() mode, with explicit dropout masks.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_train.py

Pins the *training* graph of the oracle (``forward_with_grad(..., dropout=...)``): where the reference applies its
dropouts and what gradients result.  ``nn.Dropout.forward`` is patched to multiply by pre-drawn keep masks (recorded in
call order); the one dropout that is not an ``nn.Dropout`` module - the attention-weight dropout inside
``nn.MultiheadAttention`` - is switched off for this fixture (``self_attn.dropout = 0``), its placement is fixed by
torch's documented MHA semantics rather than by wav2sleep code.  Stored: inputs' seed, labels, the masks (bit-packed),
train-mode logits, loss, the L2 norm of every parameter gradient and a few full gradient tensors.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from make_golden import SEED, build_reference, import_reference, make_inputs

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "train_dropout_cardio.npz"
SMAP = {"ABD": "ABD", "THX": "THX", "ECG": "ECG", "PPG": "PPG"}
B, S, NCLS = 2, 6, 4
FULL = ["classifier.weight", "classifier.bias", "epoch_mixer.register_tokens",
        "epoch_mixer.transformer_encoder.layers.0.norm1.weight",
        "epoch_mixer.transformer_encoder.layers.0.self_attn.in_proj_bias",
        "epoch_mixer.transformer_encoder.layers.1.linear2.bias",
        "sequence_mixer.dilated_convs.0.conv_layers.5.norm.weight",
        "sequence_mixer.dilated_convs.1.conv_layers.0.norm.bias",
        "signal_encoders.encoders.ECG.cnn.0.conv1.conv.weight", "signal_encoders.encoders.ECG.cnn.0.downsample.weight",
        "signal_encoders.encoders.ABD.cnn.5.conv3.conv.weight", "signal_encoders.encoders.PPG.linear.bias"]


def main():
    m = import_reference()
    model = build_reference(m, SMAP, NCLS).train()
    for layer in model.epoch_mixer.transformer_encoder.layers:
        layer.self_attn.dropout = 0.0
    x = make_inputs(SMAP, B, S, [("ABD", 0), ("PPG", 1)], [], seed=77)
    labels = torch.randint(0, NCLS, (B, S), generator=torch.Generator().manual_seed(3))
    labels[1, ::4] = -1
    names = {id(mod): name for name, mod in model.named_modules()}
    gen = torch.Generator().manual_seed(99)
    record = []

    def patched(self, inp):
        if self.p == 0.0 or not self.training:
            return inp
        keep = (torch.rand(inp.shape, generator=gen) >= self.p)
        record.append((names[id(self)], self.p, keep))
        return inp * keep.to(inp.dtype) / (1.0 - self.p)

    orig = torch.nn.Dropout.forward
    torch.nn.Dropout.forward = patched
    try:
        logits = model(x)
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, NCLS), labels.reshape(-1), ignore_index=-1)
        loss.backward()
    finally:
        torch.nn.Dropout.forward = orig
    out = {"logits": logits.detach().numpy(), "loss": np.float64(loss.item()), "labels": labels.numpy(),
           "input_seed": np.int64(77), "weights_seed": np.int64(SEED),
           "mask_names": np.array([r[0] for r in record]), "mask_p": np.array([r[1] for r in record])}
    for i, (_, _, keep) in enumerate(record):
        out[f"mask_{i}_shape"] = np.array(keep.shape)
        out[f"mask_{i}_bits"] = np.packbits(keep.numpy().reshape(-1))
    gn_names, gn = [], []
    for name, p in model.named_parameters():
        gn_names.append(name)
        gn.append(0.0 if p.grad is None else p.grad.norm().item())
    out["grad_norm_names"], out["grad_norms"] = np.array(gn_names), np.array(gn)
    sd = dict(model.named_parameters())
    for name in FULL:
        out["grad::" + name] = sd[name].grad.detach().numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(record), "dropout calls:", [(r[0], tuple(r[2].shape)) for r in record])


if __name__ == "__main__":
    main()
