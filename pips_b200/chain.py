"""Batched visibility-aware chaining over long videos.

The reference tracks a long clip one particle at a time (chain_demo.py:40-83, identical logic in
test_on_badja.py:65-113): run the 8-frame model from the particle's current frame, keep the window's
estimate, jump to the latest frame of the window (index 2..7) whose visibility passes a threshold that
is lowered by 0.02 per sweep, repeat -- re-running fnet on every window of every particle.

Here all particles advance together: fnet runs ONCE per frame (exact: InstanceNorm is per frame,
nets/pips.py:412), the pyramid of the whole clip stays resident, and each round is one refinement call
in which every track reads its own window through a per-track frame offset (``frame_base`` of
pips_corr_gather); the threshold sweep is evaluated for all tracks at once.  Results are the same as
the per-particle loop up to fp32 noise.
"""
from __future__ import annotations

from typing import Optional

import torch

S_WIN = 8


def _threshold_table(n: int = 64) -> torch.Tensor:
    """thr = 0.9; thr -= 0.02 ... exactly as the Python loop accumulates it (chain_demo.py:64,75)."""
    vals, thr = [], 0.9
    for _ in range(n):
        vals.append(thr)
        thr -= 0.02
    return torch.tensor(vals, dtype=torch.float32)


def pick_skip(vis: torch.Tensor, thr_table: torch.Tensor) -> torch.Tensor:
    """vis (8, n) sigmoid visibilities -> (n,) frame index in [2, 7] chosen by chain_demo.py:63-76:
    scan si = 7..2 for vis > thr, lowering thr by 0.02 whenever the scan reaches si == 1."""
    cand = vis[2:S_WIN].unsqueeze(0) > thr_table.view(-1, 1, 1).to(vis.device)          # (K, 6, n)
    level_ok = cand.any(dim=1)                                                           # (K, n)
    if not bool(level_ok[-1].all()):
        raise RuntimeError("chain: visibility below every threshold (non-finite visibilities?)")
    k_star = level_ok.float().argmax(dim=0)                                              # first passing level
    sel = cand[k_star, :, torch.arange(vis.shape[1], device=vis.device)]                 # (n, 6)
    idx = torch.arange(2, S_WIN, device=vis.device).view(1, -1)
    return (sel.long() * idx).max(dim=1).values


@torch.no_grad()
def track_chain(model, rgbs: torch.Tensor, xy0: torch.Tensor, iters: int = 6, return_rounds: bool = False,
                advance: Optional[int] = None):
    """rgbs (1, T, 3, H, W) float 0..255, xy0 (1, N, 2) start positions at frame 0 (input pixels).
    Returns trajs_e (1, T, N, 2).  Equivalent to chain_demo.py:run_model's per-particle loop.
    ``advance`` (2..7): every track moves on by that many frames per round instead of the visibility-driven
    choice -- a data-independent schedule (ceil((T-1)/advance) rounds) for throughput measurements."""
    if advance is not None and not 2 <= int(advance) <= S_WIN - 1:
        raise ValueError("advance must be in 2..7 (the reference's sweep never picks another frame, chain_demo.py:63-76)")
    B, T, C, H, W = rgbs.shape
    assert B == 1, "chained tracking follows the reference: one clip at a time"
    N = xy0.shape[1]
    dev = rgbs.device
    eng = model.engine
    stride = float(model.stride)
    fmaps = torch.cat([model.encode(rgbs[:, t0:t0 + 32]) for t0 in range(0, T, 32)], dim=1)   # (1,T,128,H8,W8)
    if not fmaps.is_contiguous() and not fmaps.reshape(T, *fmaps.shape[2:]).permute(0, 2, 3, 1).is_contiguous():
        fmaps = fmaps.contiguous()

    thr = _threshold_table().to(dev)
    traj = torch.zeros(T, N, 2, dtype=torch.float32, device=dev)
    traj[0] = xy0[0].float()
    cur = torch.zeros(N, dtype=torch.long, device=dev)
    feat: Optional[torch.Tensor] = None
    active = torch.arange(N, device=dev)
    rounds = 0
    while active.numel() > 0:
        na = active.numel()
        base = cur[active]
        start = traj[base, active]                                                   # (na, 2) current position
        coords = (start / stride).view(1, 1, na, 2).repeat(1, S_WIN, 1, 1)           # zero-velocity init, :453
        fi = None if feat is None else feat[:, active]
        preds, vis_e, ffeat = eng.refine(model, fmaps, coords, fi, iters, stride,
                                         frame_base=base.view(1, na).to(torch.int32),
                                         reuse_pyramid=rounds > 0)                  # same clip every round
        if feat is None:
            feat = ffeat                                                             # outs[3], carried (chain_demo.py:57)
        xys = preds[-1][0]                                                           # (8, na, 2)
        t_idx = base.view(1, na) + torch.arange(S_WIN, device=dev).view(S_WIN, 1)    # (8, na)
        ok = t_idx < T                                                               # S_local truncation, :61
        traj[t_idx[ok], active.view(1, na).expand(S_WIN, na)[ok]] = xys[ok]
        si = pick_skip(torch.sigmoid(vis_e[0]), thr) if advance is None else torch.full_like(base, int(advance))
        cur[active] = base + si
        active = active[cur[active] < T]
        rounds += 1
    out = traj.unsqueeze(0)
    return (out, rounds) if return_rounds else out
