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


def test_matches_the_reference_recorded_at_full_size():
    """The unmodified reference was run ONCE on CPU at exactly the bench configuration (B=4, 8x384x512, N=1024, stride 8,
    6 iterations; tests/golden/make_golden.py --cfg2 -> reference_cfg2.npz, stored subsampled).  Direct comparison of
    the CUDA path with that recording: no oracle in the loop, the north-star tolerance of 1e-3 px."""
    import os
    import numpy as np
    from tests.golden.make_golden import CFG2_CASE as c, CFG2_EVERY, case_inputs
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_cfg2.npz"))
    assert (c["B"], c["H"], c["W"], c["N"], c["iters"], c["stride"]) == (B, H, W, N, ITERS, STRIDE)
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, _ = case_inputs(c)
    model = Pips(S=S, stride=STRIDE).to(DEV).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        preds, _, vis_e, ffeat, _ = model(xys.to(DEV), rgbs.to(DEV), iters=ITERS, return_feat=True)
    p = torch.stack(preds).cpu().numpy()
    err_iter = np.abs(p[:, :, :, ::CFG2_EVERY] - gold["preds_sub"]).reshape(ITERS, -1).max(1)
    err_final = np.abs(p[-1] - gold["preds_final"]).max()
    err_vis = np.abs(vis_e.cpu().numpy() - gold["vis_e"]).max()
    err_feat = np.abs(ffeat[:, ::16].cpu().numpy() - gold["ffeat_sub"]).max()
    print(f"cfg2 full size vs reference recording: per-iter max|d trajs| px = {err_iter}, final (all particles) {err_final:.3e}, "
          f"vis_e {err_vis:.3e}, ffeat {err_feat:.3e}")
    assert err_iter.max() < 1e-3 and err_final < 1e-3
    assert err_vis < 5e-3 and err_feat < 5e-4
