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


def test_matches_the_reference_on_the_real_demo_clip():
    """BASELINE cfg 1 on the real frames demo_images/000100-000107.jpg (JPEG bytes inside the fixture, decoded with
    PIL), the recipe of demo.py:21-41: the CUDA path against the unmodified reference's recording, 1e-3 px."""
    import hashlib
    import os
    import numpy as np
    from tests.golden.make_golden import DEMO_CASE as c, demo_decode, demo_inputs
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_demo.npz"))
    raw = demo_decode([gold[f"jpeg{i}"].tobytes() for i in range(8)])
    if hashlib.sha256(raw.to(torch.uint8).numpy().tobytes()).hexdigest() != str(gold["pixels_sha256"]):
        pytest.skip("this host's JPEG decoder produces different pixels than the one the fixture was recorded with")
    rgbs, xy = demo_inputs(raw.to(DEV))                      # resize + grid on the device, like demo.py:22-36
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    model = Pips(S=S, stride=c["stride"]).to(DEV).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        preds, preds2, vis_e, ffeat, _ = model(xy.to(DEV), rgbs, iters=c["iters"], return_feat=True)
    err_iter = np.abs(torch.stack(preds).cpu().numpy() - gold["preds"]).reshape(c["iters"], -1).max(1)
    err_vis = np.abs(vis_e.cpu().numpy() - gold["vis_e"]).max()
    err_feat = np.abs(ffeat.cpu().numpy() - gold["ffeat"]).max()
    print(f"real demo clip vs reference recording: per-iter max|d trajs| px = {err_iter}, vis_e {err_vis:.3e}, ffeat {err_feat:.3e}")
    assert err_iter.max() < 1e-3 and err_vis < 5e-3 and err_feat < 5e-4
    assert len(preds2) == c["iters"] + 4


def test_matches_the_reference_at_cfg4_shape():
    """BASELINE cfg 4: B=1, 8 x 720 x 1280, N=16384, stride 8, 6 iterations in ONE call (the reference has to chunk:
    its all-pairs volume would be 10 GB per iteration); every 64th particle is the 256-particle chunk the unmodified
    reference was run on (tests/golden/make_golden.py --cfg4), 1e-3 px."""
    import os
    import numpy as np
    from tests.golden.make_golden import CFG4_CASE as c, CFG4_EVERY, case_inputs
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_cfg4.npz"))
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, _ = case_inputs(c)
    model = Pips(S=S, stride=c["stride"]).to(DEV).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        preds, _, vis_e, ffeat, _ = model(xys.to(DEV), rgbs.to(DEV), iters=c["iters"], return_feat=True)
    assert preds[-1].shape == (1, S, c["N"], 2)
    p = torch.stack(preds)[:, :, :, ::CFG4_EVERY].cpu().numpy()
    err_iter = np.abs(p - gold["preds"]).reshape(c["iters"], -1).max(1)
    err_vis = np.abs(vis_e[:, :, ::CFG4_EVERY].cpu().numpy() - gold["vis_e"]).max()
    err_feat = np.abs(ffeat[:, ::CFG4_EVERY].cpu().numpy() - gold["ffeat"]).max()
    print(f"cfg4 shape (N=16384 in one call) vs reference chunk: per-iter max|d trajs| px = {err_iter}, vis_e {err_vis:.3e}, ffeat {err_feat:.3e}")
    assert err_iter.max() < 1e-3 and err_vis < 5e-3 and err_feat < 5e-4
    # the same 256 queries alone (a different GEMM tiling and chunking) give bit-identical tracks
    with torch.no_grad():
        sub = model(xys[:, ::CFG4_EVERY].contiguous().to(DEV), rgbs.to(DEV), iters=c["iters"])[0]
    assert torch.equal(torch.stack(sub), torch.stack(preds)[:, :, :, ::CFG4_EVERY])
