"""Synthetic workloads for benchmarks and tools (no dataset or checkpoint is available offline).

Product-side helpers only: ``bench.py`` and ``tools/`` use these so that nothing on a measured path touches
``oracle/`` (the oracle has its own, independent generators for the test fixtures).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def smooth_video(B: int, S: int, H: int, W: int, seed: int = 1234, shift=(3, 2)) -> torch.Tensor:
    """(B,S,3,H,W) float clip with integer values 0..255: low-resolution uniform noise, bicubic-upsampled and
    translated ``shift`` px per frame.  Integer values are exact in bf16 and uint8."""
    g = torch.Generator().manual_seed(seed)
    pad = max(abs(shift[0]), abs(shift[1])) * S
    low = torch.rand(B, 3, (H + 2 * pad) // 16 + 2, (W + 2 * pad) // 16 + 2, generator=g)
    big = F.interpolate(low, size=(H + 2 * pad, W + 2 * pad), mode="bicubic", align_corners=False)
    frames = [big[:, :, pad + shift[1] * s: pad + shift[1] * s + H, pad + shift[0] * s: pad + shift[0] * s + W] for s in range(S)]
    return (torch.stack(frames, 1).clamp(0, 1) * 255.0).round().contiguous()


def random_queries(B: int, N: int, H: int, W: int, seed: int = 4321) -> torch.Tensor:
    """(B,N,2) query points ~ U([8,W-8] x [8,H-8]) in input pixels."""
    g = torch.Generator().manual_seed(seed)
    x = 8 + torch.rand(B, N, generator=g) * (W - 16)
    y = 8 + torch.rand(B, N, generator=g) * (H - 16)
    return torch.stack([x, y], -1).float()


def seeded_model(stride: int = 8, seed: int = 0, head_scale: float = 0.05, **kw):
    """A ``Pips`` with seeded default initialisation and the last mixer Linear damped by ``head_scale`` so that
    the random-weight model is contractive like a trained one."""
    from .pips import Pips
    torch.manual_seed(seed)
    model = Pips(S=8, stride=stride, **kw)
    with torch.no_grad():
        head = model.delta_block.to_delta[15]
        head.weight.mul_(head_scale)
        head.bias.mul_(head_scale)
    return model
