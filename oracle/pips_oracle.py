"""CPU oracle for the PIPs refinement hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-torch (CPU, fp32 or fp64) *functional* restatement of the
algorithm in the reference ``nets/pips.py`` (aharley/pips @ a05a120).  It exists
to check the CUDA product path; it is never shipped or called by the product.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import it.

Pinning: the reference has no golden vectors or unit tests of its own
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself*: ``tests/golden/make_golden.py`` imports the unmodified reference from
/root/reference, loads the state_dict produced by ``init_state_dict`` below with
``strict=True`` and records its outputs; ``tests/test_oracle_golden.py`` replays
them against this file.  Recorded configurations: five small inference cases, the
demo size, a supervised / ``is_train`` call with its score maps, a chained track,
BASELINE cfg 2 at full size, the cfg 4 shape (8 x 720 x 1280; the 256-particle chunk
the reference can hold) and the real demo frames ``demo_images/000100-000107.jpg``.

Every function cites the reference lines it restates (paths relative to the
reference root).  Model state is a flat ``dict[str, Tensor]`` with exactly the
reference's ``state_dict`` keys.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

LATENT = 128          # nets/pips.py:408
CORR_LEVELS = 4       # nets/pips.py:409
CORR_RADIUS = 3       # nets/pips.py:410
MIXER_DIM = 512       # nets/pips.py:298
MIXER_DEPTH = 12      # nets/pips.py:300
KITCHEN = CORR_LEVELS * (2 * CORR_RADIUS + 1) ** 2 + LATENT + 64 * 3 + 3   # 519, nets/pips.py:289


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------

def state_dict_spec(S: int = 8) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names and shapes of every tensor in ``nets.pips.Pips(S).state_dict()``
    (nets/pips.py:131-244 for fnet, :111-123 for the mixer, :416-426 heads)."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, cout, cin, k):
        spec.append((name + ".weight", (cout, cin, k, k)))
        spec.append((name + ".bias", (cout,)))

    conv("fnet.conv1", 64, 3, 7)
    cin = 64
    for li, (dim, stride) in enumerate([(64, 1), (96, 2), (128, 2), (128, 2)], start=1):
        for bi in range(2):
            pre = f"fnet.layer{li}.{bi}"
            conv(pre + ".conv1", dim, cin if bi == 0 else dim, 3)
            conv(pre + ".conv2", dim, dim, 3)
            if bi == 0 and stride != 1:
                conv(pre + ".downsample.0", dim, cin, 1)
        cin = dim
    conv("fnet.conv2", 256, 128 + 128 + 96 + 64, 3)
    conv("fnet.conv3", LATENT, 256, 1)

    td = "delta_block.to_delta"
    spec += [(f"{td}.0.weight", (MIXER_DIM, KITCHEN)), (f"{td}.0.bias", (MIXER_DIM,))]
    for l in range(1, MIXER_DEPTH + 1):
        spec += [
            (f"{td}.{l}.0.fn.0.weight", (4 * S, S, 1)), (f"{td}.{l}.0.fn.0.bias", (4 * S,)),
            (f"{td}.{l}.0.fn.3.weight", (S, 4 * S, 1)), (f"{td}.{l}.0.fn.3.bias", (S,)),
            (f"{td}.{l}.0.norm.weight", (MIXER_DIM,)), (f"{td}.{l}.0.norm.bias", (MIXER_DIM,)),
            (f"{td}.{l}.1.fn.0.weight", (4 * MIXER_DIM, MIXER_DIM)), (f"{td}.{l}.1.fn.0.bias", (4 * MIXER_DIM,)),
            (f"{td}.{l}.1.fn.3.weight", (MIXER_DIM, 4 * MIXER_DIM)), (f"{td}.{l}.1.fn.3.bias", (MIXER_DIM,)),
            (f"{td}.{l}.1.norm.weight", (MIXER_DIM,)), (f"{td}.{l}.1.norm.bias", (MIXER_DIM,)),
        ]
    spec += [(f"{td}.13.weight", (MIXER_DIM,)), (f"{td}.13.bias", (MIXER_DIM,)),
             (f"{td}.15.weight", (S * (LATENT + 2), MIXER_DIM)), (f"{td}.15.bias", (S * (LATENT + 2),)),
             ("norm.weight", (LATENT,)), ("norm.bias", (LATENT,)),
             ("ffeat_updater.0.weight", (LATENT, LATENT)), ("ffeat_updater.0.bias", (LATENT,)),
             ("vis_predictor.0.weight", (1, LATENT)), ("vis_predictor.0.bias", (1,))]
    return spec


def init_state_dict(seed: int = 0, head_scale: float = 0.05, S: int = 8, norm_jitter: float = 0.1) -> SD:
    """Seeded random weights with the reference's names/shapes.

    No trained checkpoint is available offline.  The
    distributions follow torch's defaults (U(+-1/sqrt(fan_in)) for Linear/Conv1d,
    kaiming-normal fan_out for fnet convs as nets/pips.py:229-231); norm affine
    parameters get a small jitter so a swapped weight/bias cannot hide.
    ``head_scale`` damps the last mixer Linear, which makes the random-weight
    model contractive like a trained one; without it the
    reference disagrees with *itself* by ~1 px after 6 iterations.
    """
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in state_dict_spec(S):
        if name.startswith("fnet.") and name.endswith(".weight"):
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif name.startswith("fnet."):
            fan_in = None  # bias: torch default U(+-1/sqrt(fan_in)) of the matching conv
            w = sd[name[:-4] + "weight"]
            fan_in = w.shape[1] * w.shape[2] * w.shape[3]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        elif ".norm." in name or name.startswith("norm.") or ".13." in name:
            base = 1.0 if name.endswith("weight") else 0.0
            t = base + norm_jitter * (torch.rand(shape, generator=g) * 2 - 1)
        elif name.endswith(".weight"):
            fan_in = shape[1]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            w = sd[name[:-4] + "weight"]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(w.shape[1])
        sd[name] = t.float()
    sd["delta_block.to_delta.15.weight"] *= head_scale
    sd["delta_block.to_delta.15.bias"] *= head_scale
    return sd


def smooth_video(B: int, S: int, H: int, W: int, seed: int = 1234, shift=(3, 2)) -> Tensor:
    """Synthetic clip used by fixtures and the bench: low-res
    uniform noise, bicubic-upsampled, translated ``shift`` px/frame, rounded to
    integers in [0,255] (so it is exactly representable in bf16 and uint8)."""
    g = torch.Generator().manual_seed(seed)
    pad = max(abs(shift[0]), abs(shift[1])) * S
    low = torch.rand(B, 3, (H + 2 * pad) // 16 + 2, (W + 2 * pad) // 16 + 2, generator=g)
    big = F.interpolate(low, size=(H + 2 * pad, W + 2 * pad), mode="bicubic", align_corners=False)
    frames = []
    for s in range(S):
        y0, x0 = pad + shift[1] * s, pad + shift[0] * s
        frames.append(big[:, :, y0:y0 + H, x0:x0 + W])
    vid = torch.stack(frames, 1).clamp(0, 1) * 255.0
    return vid.round().contiguous()


def random_queries(B: int, N: int, H: int, W: int, seed: int = 4321) -> Tensor:
    """xys ~ U([8,W-8] x [8,H-8])."""
    g = torch.Generator().manual_seed(seed)
    x = 8 + torch.rand(B, N, generator=g) * (W - 16)
    y = 8 + torch.rand(B, N, generator=g) * (H - 16)
    return torch.stack([x, y], -1).float()


# ----------------------------------------------------------------------------
# fnet (BasicEncoder, instance norm)  --  upstream of the hot path, a13
# ----------------------------------------------------------------------------

def _inorm(x: Tensor) -> Tensor:
    # nn.InstanceNorm2d defaults: affine=False, eps=1e-5, no running stats
    return F.instance_norm(x, eps=1e-5)


def _res_block(sd: SD, pre: str, x: Tensor, stride: int) -> Tensor:
    """nets/pips.py:173-181 (ResidualBlock.forward, norm_fn='instance')."""
    y = F.relu(_inorm(F.conv2d(x, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], stride=stride, padding=1)))
    y = F.relu(_inorm(F.conv2d(y, sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], padding=1)))
    if stride != 1:
        x = _inorm(F.conv2d(x, sd[pre + ".downsample.0.weight"], sd[pre + ".downsample.0.bias"], stride=stride))
    return F.relu(x + y)


def fnet(sd: SD, x: Tensor, stride: int) -> Tensor:
    """nets/pips.py:247-281 (BasicEncoder.forward, non-shallow branch)."""
    _, _, H, W = x.shape
    x = F.relu(_inorm(F.conv2d(x, sd["fnet.conv1.weight"], sd["fnet.conv1.bias"], stride=2, padding=3)))
    outs = []
    for li, st in zip((1, 2, 3, 4), (1, 2, 2, 2)):
        x = _res_block(sd, f"fnet.layer{li}.0", x, st)
        x = _res_block(sd, f"fnet.layer{li}.1", x, 1)
        outs.append(F.interpolate(x, (H // stride, W // stride), mode="bilinear", align_corners=True))
    x = F.conv2d(torch.cat(outs, 1), sd["fnet.conv2.weight"], sd["fnet.conv2.bias"], padding=1)
    x = F.relu(_inorm(x))
    return F.conv2d(x, sd["fnet.conv3.weight"], sd["fnet.conv3.bias"])


# ----------------------------------------------------------------------------
# correlation
# ----------------------------------------------------------------------------

def build_pyramid(fmaps: Tensor, levels: int = CORR_LEVELS) -> List[Tensor]:
    """nets/pips.py:346-352: 2x2 average pooling, floor on odd sizes.  fmaps (B,S,C,H,W)."""
    B, S, C, H, W = fmaps.shape
    pyr = [fmaps]
    for _ in range(levels - 1):
        f = F.avg_pool2d(pyr[-1].reshape(B * S, C, *pyr[-1].shape[-2:]), 2, stride=2)
        pyr.append(f.reshape(B, S, C, *f.shape[-2:]))
    return pyr


def corr_allpairs(pyr: List[Tensor], targets: Tensor) -> List[Tensor]:
    """nets/pips.py:384-398 (CorrBlock.corr): all-pairs volume per level, / sqrt(C)."""
    B, S, N, C = targets.shape
    vols = []
    for f in pyr:
        H, W = f.shape[-2:]
        v = torch.matmul(targets, f.reshape(B, S, C, H * W)).reshape(B, S, N, H, W)
        vols.append(v / torch.sqrt(torch.tensor(C).float()).to(v.dtype))
    return vols


def sample_allpairs(vols: List[Tensor], coords: Tensor, r: int = CORR_RADIUS) -> Tensor:
    """nets/pips.py:355-382 (CorrBlock.sample) + :313-328 (bilinear_sampler).
    Channel k = a*(2r+1)+b of a level samples x = cx + (a-r), y = cy + (b-r)
    (the meshgrid(dy,dx) 'ij' quirk)."""
    B, S, N, _ = coords.shape
    out = []
    d = torch.linspace(-r, r, 2 * r + 1, dtype=coords.dtype)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)          # (2r+1,2r+1,2): [a,b] -> (d[a], d[b])
    for i, v in enumerate(vols):
        H, W = v.shape[-2:]
        cl = coords.reshape(B * S * N, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        xg = 2 * cl[..., 0:1] / (W - 1) - 1
        yg = 2 * cl[..., 1:2] / (H - 1) - 1
        smp = F.grid_sample(v.reshape(B * S * N, 1, H, W), torch.cat([xg, yg], -1), align_corners=True)
        out.append(smp.view(B, S, N, -1))
    return torch.cat(out, -1).contiguous()


def corr_local(pyr: List[Tensor], targets: Tensor, coords: Tensor, r: int = CORR_RADIUS) -> Tensor:
    """Same quantity as ``sample_allpairs(corr_allpairs(...))`` computed without the
    volume: by linearity, bilinear-sampling the dot-product map equals the dot
    product with bilinearly sampled features; the 49 taps of a level share one
    fractional offset so they touch an 8x8 pixel footprint with zero padding
    outside the map (grid_sample default padding_mode='zeros', nets/pips.py:322).
    Used for large N where the volume does not fit; scales O(N)."""
    B, S, N, C = targets.shape
    out = []
    n = 2 * r + 2
    for i, f in enumerate(pyr):
        H, W = f.shape[-2:]
        c = coords / 2 ** i
        x0 = torch.floor(c[..., 0])
        y0 = torch.floor(c[..., 1])
        fx = (c[..., 0] - x0).unsqueeze(-1).unsqueeze(-1)
        fy = (c[..., 1] - y0).unsqueeze(-1).unsqueeze(-1)
        ix = x0.long().unsqueeze(-1) + torch.arange(-r, r + 2)                  # (B,S,N,8)
        iy = y0.long().unsqueeze(-1) + torch.arange(-r, r + 2)
        okx = (ix >= 0) & (ix < W)
        oky = (iy >= 0) & (iy < H)
        flat = f.permute(0, 1, 3, 4, 2).reshape(B, S, H * W, C)
        lin = (iy.clamp(0, H - 1).unsqueeze(-1) * W + ix.clamp(0, W - 1).unsqueeze(-2))   # (B,S,N,8y,8x)
        g = torch.gather(flat, 2, lin.reshape(B, S, N * n * n, 1).expand(-1, -1, -1, C))
        g = g.reshape(B, S, N, n, n, C)
        dots = (g * targets.reshape(B, S, N, 1, 1, C)).sum(-1) / math.sqrt(C)             # [y, x]
        dots = dots * (oky.unsqueeze(-1) & okx.unsqueeze(-2)).to(dots.dtype)
        blend = ((1 - fy) * (1 - fx) * dots[..., :-1, :-1] + (1 - fy) * fx * dots[..., :-1, 1:]
                 + fy * (1 - fx) * dots[..., 1:, :-1] + fy * fx * dots[..., 1:, 1:])      # [b(y), a(x)]
        out.append(blend.transpose(-1, -2).reshape(B, S, N, -1))                         # k = a*7+b
    return torch.cat(out, -1).contiguous()


def bilinear_sample2d(im: Tensor, x: Tensor, y: Tensor) -> Tensor:
    """utils/samp.py:5-78: indices clamped to the map, weights NOT clamped.
    im (B,C,H,W); x,y (B,N) -> (B,N,C)."""
    B, C, H, W = im.shape
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    x1, y1 = x0 + 1, y0 + 1
    xi0, xi1 = x0.long().clamp(0, W - 1), x1.long().clamp(0, W - 1)
    yi0, yi1 = y0.long().clamp(0, H - 1), y1.long().clamp(0, H - 1)
    flat = im.permute(0, 2, 3, 1).reshape(B, H * W, C)

    def at(yi, xi):
        return torch.gather(flat, 1, (yi * W + xi).unsqueeze(-1).expand(-1, -1, C))

    w00 = ((x1 - x) * (y1 - y)).unsqueeze(-1)
    w01 = ((x - x0) * (y1 - y)).unsqueeze(-1)
    w10 = ((x1 - x) * (y - y0)).unsqueeze(-1)
    w11 = ((x - x0) * (y - y0)).unsqueeze(-1)
    return w00 * at(yi0, xi0) + w01 * at(yi0, xi1) + w10 * at(yi1, xi0) + w11 * at(yi1, xi1)


# ----------------------------------------------------------------------------
# delta block
# ----------------------------------------------------------------------------

def embedding3d(xyz: Tensor, C: int = 64) -> Tensor:
    """utils/misc.py:44-69 (get_3d_embedding, cat_coords=True).  xyz (R,S,3) -> (R,S,3C+3)."""
    div = (torch.arange(0, C, 2, dtype=torch.float32) * (1000.0 / C)).to(xyz.dtype).reshape(1, 1, C // 2)
    parts = []
    for a in range(3):
        v = xyz[:, :, a:a + 1] * div
        pe = torch.stack([torch.sin(v), torch.cos(v)], -1).reshape(*xyz.shape[:2], C)   # even=sin, odd=cos
        parts.append(pe)
    return torch.cat(parts + [xyz], 2)


def _ln(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def mixer(sd: SD, x: Tensor, trace: Optional[dict] = None) -> Tensor:
    """nets/pips.py:111-123 (MLPMixer) with :93-109.  x (R,S,519) -> (R, S*130)."""
    td = "delta_block.to_delta"
    x = F.linear(x, sd[f"{td}.0.weight"], sd[f"{td}.0.bias"])
    if trace is not None:
        trace["x0"] = x.clone()
    for l in range(1, MIXER_DEPTH + 1):
        p = f"{td}.{l}"
        # token mixing: Conv1d(k=1) over the S axis (channels-first => dim 1 is S)
        y = _ln(x, sd[p + ".0.norm.weight"], sd[p + ".0.norm.bias"])
        h = torch.einsum("js,rsc->rjc", sd[p + ".0.fn.0.weight"][:, :, 0], y) + sd[p + ".0.fn.0.bias"].view(1, -1, 1)
        h = F.gelu(h)
        y = torch.einsum("sj,rjc->rsc", sd[p + ".0.fn.3.weight"][:, :, 0], h) + sd[p + ".0.fn.3.bias"].view(1, -1, 1)
        x = x + y
        # channel mixing
        y = _ln(x, sd[p + ".1.norm.weight"], sd[p + ".1.norm.bias"])
        y = F.linear(F.gelu(F.linear(y, sd[p + ".1.fn.0.weight"], sd[p + ".1.fn.0.bias"])),
                     sd[p + ".1.fn.3.weight"], sd[p + ".1.fn.3.bias"])
        x = x + y
        if trace is not None and l in (1, MIXER_DEPTH):
            trace[f"x{l}"] = x.clone()
    x = _ln(x, sd[f"{td}.13.weight"], sd[f"{td}.13.bias"]).mean(1)              # Reduce('b n c -> b c','mean')
    return F.linear(x, sd[f"{td}.15.weight"], sd[f"{td}.15.bias"])


def delta_block(sd: SD, fhid: Tensor, fcorr: Tensor, flow: Tensor, trace: Optional[dict] = None) -> Tensor:
    """nets/pips.py:304-311 (DeltaBlock.forward).  All (R,S,*) -> (R,S,130)."""
    R, S, _ = flow.shape
    x = torch.cat([fhid, fcorr, embedding3d(flow, 64)], 2)
    if trace is not None:
        trace["mixer_in"] = x.clone()
    return mixer(sd, x, trace).reshape(R, S, LATENT + 2)


def times_axis(S: int, dtype=torch.float32) -> Tensor:
    """nets/pips.py:519: linspace(0, S, S) = [0, S/(S-1), ..., S]  (not 0..S-1)."""
    return torch.linspace(0, S, S, dtype=torch.float32).to(dtype)


def refine_iter(sd: SD, pyr: List[Tensor], coords: Tensor, ffeats: Tensor, coords0: Tensor,
                allpairs: bool = False, is_train: bool = False, trace: Optional[dict] = None
                ) -> Tuple[Tensor, Tensor]:
    """One pass of the loop body nets/pips.py:499-539 (without the dead `fcp`
    heat-map of :504-511).  coords (B,S,N,2) in feature-map px, ffeats (B,S,N,128)."""
    B, S, N, _ = coords.shape
    if allpairs:
        fcorrs = sample_allpairs(corr_allpairs(pyr, ffeats), coords)                # :502, :513
    else:
        fcorrs = corr_local(pyr, ffeats, coords)
    if trace is not None:
        trace["fcorrs"] = fcorrs.clone()
    LRR = fcorrs.shape[3]
    fcorrs_ = fcorrs.permute(0, 2, 1, 3).reshape(B * N, S, LRR)                        # :517
    flows_ = (coords - coords[:, 0:1]).permute(0, 2, 1, 3).reshape(B * N, S, 2)        # :518
    times_ = times_axis(S, coords.dtype).reshape(1, S, 1).repeat(B * N, 1, 1)          # :519
    flows_ = torch.cat([flows_, times_], 2)
    ffeats_ = ffeats.permute(0, 2, 1, 3).reshape(B * N, S, LATENT)                     # :522
    delta = delta_block(sd, ffeats_, fcorrs_, flows_, trace)                           # :524
    if trace is not None:
        trace["delta"] = delta.clone()
    dcoords, dfeats = delta[:, :, :2], delta[:, :, 2:]
    ffeats_ = ffeats_.reshape(B * N * S, LATENT)
    dfeats = dfeats.reshape(B * N * S, LATENT)
    g = F.group_norm(dfeats, 1, sd["norm.weight"], sd["norm.bias"], 1e-5)              # :416, :530
    ffeats_ = F.gelu(F.linear(g, sd["ffeat_updater.0.weight"], sd["ffeat_updater.0.bias"])) + ffeats_
    ffeats = ffeats_.reshape(B, N, S, LATENT).permute(0, 2, 1, 3)                      # :531
    coords = coords + dcoords.reshape(B, N, S, 2).permute(0, 2, 1, 3)                  # :533
    if not is_train:
        coords = coords.clone()
        coords[:, 0] = coords0[:, 0]                                                   # :535-536
    return coords, ffeats


def heatmap_fcp(vols: List[Tensor], H8: int, W8: int) -> Tensor:
    """nets/pips.py:504-511: the dense score map (dead at inference, timed only for
    the faithful CPU baseline)."""
    B, S, N = vols[0].shape[:3]
    fcp = torch.zeros(B, S, N, H8, W8, dtype=vols[0].dtype)
    for v in vols:
        h, w = v.shape[-2:]
        fcp = fcp + F.interpolate(v.reshape(B * S, N, h, w), (H8, W8), mode="bilinear",
                                  align_corners=True).reshape(B, S, N, H8, W8)
    return fcp


def forward(sd: SD, xys: Tensor, rgbs: Tensor, iters: int = 3, stride: int = 8,
            coords_init: Optional[Tensor] = None, feat_init: Optional[Tensor] = None,
            fmaps: Optional[Tensor] = None, allpairs: bool = False, faithful_dead_work: bool = False,
            return_feat: bool = False, traces: Optional[list] = None, dtype=torch.float32):
    """nets/pips.py:428-611 (Pips.forward) for the inference case (trajs_g=None,
    sw=None).  Returns (coord_predictions, coord_predictions2, vis_e[, ffeat], None)."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    B, N, D = xys.shape
    assert D == 2
    _, S, C, H, W = rgbs.shape
    if fmaps is None:
        x = 2 * (rgbs.to(dtype) / 255.0) - 1.0                                         # :436
        fmaps = fnet(sd, x.reshape(B * S, C, H, W), stride).reshape(B, S, LATENT, H // stride, W // stride)
    fmaps = fmaps.to(dtype)
    xys_ = xys.to(dtype) / float(stride)                                               # :450
    if coords_init is None:
        coords = xys_.reshape(B, 1, N, 2).repeat(1, S, 1, 1)                           # :453
    else:
        coords = coords_init.to(dtype) / stride                                        # :455
    pyr = build_pyramid(fmaps)                                                         # :459
    if feat_init is None:
        ffeat = bilinear_sample2d(fmaps[:, 0], coords[:, 0, :, 0], coords[:, 0, :, 1]) # :463
    else:
        ffeat = feat_init.to(dtype)
    ffeats = ffeat.unsqueeze(1).repeat(1, S, 1, 1)                                     # :466
    coords0 = coords.clone()
    preds, preds2 = [], [coords * stride, coords * stride]                             # :474-475
    for _ in range(iters):
        tr = {} if traces is not None else None
        if tr is not None:
            tr["coords_in"], tr["ffeats_in"] = coords.clone(), ffeats.clone()
        if faithful_dead_work:
            heatmap_fcp(corr_allpairs(pyr, ffeats), H // stride, W // stride)
        coords, ffeats = refine_iter(sd, pyr, coords, ffeats, coords0, allpairs=allpairs, trace=tr)
        if tr is not None:
            tr["coords_out"], tr["ffeats_out"] = coords.clone(), ffeats.clone()
            traces.append(tr)
        preds.append(coords * stride)                                                  # :538
        preds2.append(coords * stride)
    vis_e = F.linear(ffeats.reshape(B * S * N, LATENT), sd["vis_predictor.0.weight"],
                     sd["vis_predictor.0.bias"]).reshape(B, S, N)                      # :559
    preds2 += [coords * stride, coords * stride]                                       # :562-563
    if return_feat:
        return preds, preds2, vis_e, ffeat, None
    return preds, preds2, vis_e, None


def chain_track(forward_fn, rgbs: Tensor, xy0: Tensor, iters: int = 6):
    """chain_demo.py:40-83 (run_model's per-particle loop; test_on_badja.py:65-113 is the same logic).
    ``forward_fn(xys (1,1,2), rgb_seq (1,8,3,H,W), feat_init) -> (preds, vis_e, ffeat)`` is one 8-frame
    model call with return_feat=True.  rgbs (1,T,3,H,W), xy0 (1,N,2) -> trajs (1,T,N,2), skips per particle."""
    B, T = rgbs.shape[:2]
    N = xy0.shape[1]
    trajs = torch.zeros(B, T, N, 2, dtype=torch.float32)
    skips = []
    for n in range(N):
        cur, done, feat_init = 0, False, None
        traj = torch.zeros(B, T, 2, dtype=torch.float32)
        traj[:, 0] = xy0[:, n]
        hist = []
        while not done:
            end = cur + 8
            seq = rgbs[:, cur:end]
            s_local = seq.shape[1]
            seq = torch.cat([seq, seq[:, -1].unsqueeze(1).repeat(1, 8 - s_local, 1, 1, 1)], dim=1)    # :50-52
            preds, vis, feat_init = forward_fn(traj[:, cur].reshape(1, -1, 2), seq, feat_init)      # :54-57
            vis = torch.sigmoid(vis)                                                                  # :59
            traj[:, cur:end] = preds[-1].reshape(1, 8, 2)[:, :s_local]                                # :60-61
            thr, si = 0.9, 7
            while True:                                                                               # :63-76
                if vis[0, si] > thr:
                    break
                si -= 1
                if si == 1:
                    thr -= 0.02
                    si = 7
            hist.append(si)
            cur += si                                                                                 # :79
            done = cur >= T
        trajs[:, :, n] = traj
        skips.append(hist)
    return trajs, skips
