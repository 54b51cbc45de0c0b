"""``SleepPPGNet`` baseline: drop-in mirror of ``wav2sleep.models.ppgnet.SleepPPGNet``
(/root/reference/src/wav2sleep/models/ppgnet.py:19-126) whose forward runs on the general fp32 CUDA kernels
(csrc/general.cuh through general.py): eight ``ConvBlock1D`` (batch norm, leaky ReLU) -> time-distributed
``Linear(1024 -> F)`` -> two ``DilatedConvBlock`` -> classifier.  Same constructor keywords, parameter names and default
initialisation as the reference; inference only (BatchNorm uses its running statistics), not tuned.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .model import ConvBlock1D, DilatedConvBlock, get_activation

__all__ = ("SleepPPGNet",)


class WindowEncoder(nn.Module):
    """ppgnet.py:77-101"""

    CHANNELS = [16, 16, 32, 32, 64, 64, 128, 256]

    def __init__(self, activation: str = "leaky", norm: str = "batch") -> None:
        super().__init__()
        blocks, cin = [], 1
        for cout in self.CHANNELS:
            blocks.append(ConvBlock1D(cin, cout, activation=activation, norm=norm))
            cin = cout
        self.model = nn.Sequential(*blocks)


class DenseBlock(nn.Module):
    """ppgnet.py:104-126"""

    def __init__(self, in_dim: int = 1024, out_dim: int = 128, activation: str = "leaky") -> None:
        super().__init__()
        self.linear = nn.Linear(in_dim, out_dim)
        self.activation = get_activation(activation)
        self.activation_name = activation


class SleepPPGNet(nn.Module):
    INPUT_LENGTH: int = 1228800  # ppgnet.py:20

    def __init__(self, n_classes: int = 4, feature_dim: int = 128, dropout: float = 0.2, activation: str = "leaky",
                 norm: str = "batch") -> None:
        super().__init__()
        self.feature_dim = feature_dim
        self.conv_block = WindowEncoder(activation=activation, norm=norm)
        self.dense = DenseBlock(in_dim=1024, out_dim=feature_dim)
        self.dilated_convs = nn.Sequential(*[
            DilatedConvBlock(feature_dim=feature_dim, dropout=dropout, activation=activation, norm=norm) for _ in range(2)])
        self.classifier = nn.Linear(in_features=feature_dim, out_features=n_classes)
        self._general = None

    def _engine(self):
        from .general import GeneralEngine  # deferred: loads the CUDA library
        if self._general is None:
            object.__setattr__(self, "_general", GeneralEngine(self))
        return self._general

    @torch.no_grad()
    def encode(self, x_BT: Tensor) -> Tensor:
        """[N, 1228800] -> features [N, 1200, F]   (ppgnet.py:42-59)."""
        if x_BT.dim() != 2 or x_BT.size(1) != self.INPUT_LENGTH:
            raise ValueError(f"Input tensor had unexpected shape: {x_BT.size()}")
        if not x_BT.is_cuda:
            raise RuntimeError("wav2sleep_b200 runs on CUDA (sm_100a) only: move inputs with .to('cuda'); "
                               "there is no CPU fallback")
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("wav2sleep_b200: SleepPPGNet runs the general inference kernels only")
        eng = self._engine()
        B = x_BT.size(0)
        with torch.cuda.device(x_BT.device):
            a = x_BT.detach().to(torch.float32).contiguous().view(B, self.INPUT_LENGTH, 1)
            for blk in self.conv_block.model:
                a = eng.conv_block(a, blk)
            # [B, 4800, 256] channels-last == transpose(-1, -2) of the reference tensor; reshape(-1, 1200, 1024)
            rows = a.reshape(B * 1200, 1024)
            z = eng.affine_act(eng.linear(rows, self.dense.linear).view(B, 1200, -1), self.dense.activation_name)
            return eng.dilated_blocks(z, self.dilated_convs)

    def forward(self, x_BT: Tensor) -> Tensor:
        """[N, 1228800] -> logits [N, 1200, n_classes]   (ppgnet.py:61-74)."""
        feat = self.encode(x_BT)
        B, S, Fd = feat.shape
        with torch.cuda.device(feat.device):
            return self._engine().linear(feat.reshape(B * S, Fd), self.classifier).view(B, S, -1)
