"""Feature encoder (`fnet`) feeding the refinement loop.

Upstream of the hot path: it stays on PyTorch/cuDNN.  The module tree
reproduces the parameter names of the reference's ``BasicEncoder(norm_fn='instance')``
(nets/pips.py:183-281, residual blocks :131-181) so that ``saverloader.load`` /
``load_state_dict`` of a reference checkpoint fills it: conv weights and biases only (InstanceNorm has
no parameters).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _inorm(x: torch.Tensor) -> torch.Tensor:
    return F.instance_norm(x, eps=1e-5)


def _tf32_hi(x: torch.Tensor) -> torch.Tensor:
    """x with the 13 low mantissa bits cleared: exactly representable in TF32."""
    return (x.view(torch.int32) & -8192).view(torch.float32)


def conv2d_3xtf32(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """fp32-accurate convolution on the TF32 tensor-core path: with x = x_hi + x_lo and w = w_hi + w_lo
    (hi = TF32-exact part), conv(x, w) ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, evaluated as ONE TF32
    convolution over the channel-concatenated operands [x_hi | x_lo | x_hi] x [w_hi | w_hi | w_lo] so the
    three terms are summed in the fp32 accumulator.  Dropped term and re-rounding of the lo parts are
    O(2^-21) relative -- the same class as fp32 rounding itself."""
    w = conv.weight
    key = (w.data_ptr(), w._version)
    if getattr(conv, "_w3_key", None) != key:
        wd = w.detach()
        hi = _tf32_hi(wd)
        conv._w3 = torch.cat([hi, hi, wd - hi], dim=1).contiguous()
        conv._w3_key = key
    xh = _tf32_hi(x)
    return F.conv2d(torch.cat([xh, x - xh, xh], dim=1), conv._w3, conv.bias, conv.stride, conv.padding)


class ResBlock(nn.Module):
    """Two 3x3 convs with instance norm; strided 1x1 projection on the skip when downsampling
    (nets/pips.py:131-181)."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride)) if stride != 1 else None

    def forward(self, x, conv=lambda m, t: m(t)):
        y = F.relu(_inorm(conv(self.conv1, x)))
        y = F.relu(_inorm(conv(self.conv2, y)))
        if self.downsample is not None:
            x = _inorm(conv(self.downsample[0], x))
        return F.relu(x + y)


class Encoder(nn.Module):
    def __init__(self, output_dim: int = 128, stride: int = 8):
        super().__init__()
        self.stride = stride
        # 'plain': cuDNN as configured by the caller.  'x3': every convolution through conv2d_3xtf32
        # (needs torch.backends.cudnn.allow_tf32 = True; set by Pips.encode).
        self.mode = "plain"
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        widths = [(64, 64, 1), (64, 96, 2), (96, 128, 2), (128, 128, 2)]
        for i, (cin, cout, st) in enumerate(widths, start=1):
            setattr(self, f"layer{i}", nn.Sequential(ResBlock(cin, cout, st), ResBlock(cout, cout, 1)))
        self.conv2 = nn.Conv2d(64 + 96 + 128 + 128, output_dim * 2, 3, padding=1)
        self.conv3 = nn.Conv2d(output_dim * 2, output_dim, 1)
        for m in self.modules():                     # nets/pips.py:229-231
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        H, W = x.shape[-2:]
        size = (H // self.stride, W // self.stride)
        conv = conv2d_3xtf32 if self.mode == "x3" else (lambda m, t: m(t))
        x = F.relu(_inorm(conv(self.conv1, x)))
        taps = []
        for i in range(1, 5):
            for blk in getattr(self, f"layer{i}"):
                x = blk(x, conv)
            taps.append(F.interpolate(x, size, mode="bilinear", align_corners=True))
        x = F.relu(_inorm(conv(self.conv2, torch.cat(taps, dim=1))))
        return conv(self.conv3, x)
