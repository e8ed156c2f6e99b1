"""BASELINE full-size configuration (cfg 2: B=4, 8x384x512, N=1024, iters=6) through the public module API:
size-independent properties that need no oracle run (the oracle would take minutes at this size)."""
import pytest
import torch

from oracle import pips_oracle as po          # input / weight generator only
from pips_b200 import Pips

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, S, H, W, N, ITERS, STRIDE = 4, 8, 384, 512, 1024, 6, 8


@pytest.fixture(scope="module")
def full_run():
    sd = po.init_state_dict(seed=0, head_scale=0.05)
    rgbs = po.smooth_video(B, S, H, W, seed=1234).to(torch.bfloat16).to(DEV)
    xys = po.random_queries(B, N, H, W, seed=4321).to(DEV)
    model = Pips(S=S, stride=STRIDE).to(DEV).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = model(xys, rgbs, iters=ITERS)
    torch.cuda.synchronize()
    return model, rgbs, xys, out


def test_shapes_and_finiteness(full_run):
    _, _, xys, (preds, preds2, vis_e, losses) = full_run
    assert losses is None and len(preds) == ITERS and len(preds2) == ITERS + 4
    assert all(p.shape == (B, S, N, 2) and p.dtype == torch.float32 for p in preds)
    assert vis_e.shape == (B, S, N)
    assert all(bool(torch.isfinite(p).all()) for p in preds) and bool(torch.isfinite(vis_e).all())
    # the damped random model moves points by a bounded amount
    assert float((preds[-1] - xys[:, None]).abs().max()) < 64.0


def test_frame0_is_locked_to_the_query_every_iteration(full_run):
    """nets/pips.py:535-536: coords[:,0] = coords_bak[:,0] at inference -> exact equality with xys/stride*stride."""
    _, _, xys, (preds, _, _, _) = full_run
    want = (xys / float(STRIDE)) * float(STRIDE)
    for p in preds:
        assert torch.equal(p[:, 0], want)


def test_deterministic_replay(full_run):
    model, rgbs, xys, (preds, _, vis_e, _) = full_run
    with torch.no_grad():
        preds_b, _, vis_b, _ = model(xys, rgbs, iters=ITERS)
    assert all(torch.equal(a, b) for a, b in zip(preds, preds_b)) and torch.equal(vis_e, vis_b)


def test_particles_are_independent(full_run):
    """Tracking a subset of the queries gives bit-identical tracks (the mixer never mixes across N,
    nets/pips.py:517-524) -- also exercises a different GEMM tiling / row placement."""
    model, rgbs, xys, (preds, _, vis_e, _) = full_run
    idx = torch.arange(37, 37 + 200, device=DEV)
    with torch.no_grad():
        sub, _, vis_s, _ = model(xys[:, idx].contiguous(), rgbs, iters=ITERS)
    assert torch.equal(sub[-1], preds[-1][:, :, idx])
    assert torch.equal(vis_s, vis_e[:, :, idx])


def test_input_dtype_does_not_matter_for_integer_video(full_run):
    """uint8-valued frames are exact in bf16 and fp32: identical results whichever dtype the caller passes."""
    model, rgbs, xys, (preds, _, _, _) = full_run
    with torch.no_grad():
        p32, _, _, _ = model(xys[:, :64].contiguous(), rgbs.float(), iters=2)
        p16, _, _, _ = model(xys[:, :64].contiguous(), rgbs, iters=2)
    assert torch.equal(p32[-1], p16[-1])
