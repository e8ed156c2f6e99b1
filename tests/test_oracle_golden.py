"""Pins oracle/pips_oracle.py against outputs of the unmodified reference
(tests/golden/reference_outputs.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import pips_oracle as po
from tests.golden.make_golden import CASES, CPU_CASES, case_inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.npz"))

# fp32 re-association noise of the reference against itself (1 vs 8 threads) on the
# damped fixtures is <= 2e-4 px after 6 iterations.
TOL_PX = 1e-3


@pytest.mark.parametrize("name", list(CASES) + list(CPU_CASES))
@pytest.mark.parametrize("allpairs", [True, False])
def test_oracle_matches_reference(name, allpairs):
    c = CASES.get(name) or CPU_CASES[name]
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, extra = case_inputs(c)
    with torch.no_grad():
        preds, preds2, vis_e, ffeat, losses = po.forward(
            sd, xys, rgbs, iters=c["iters"], stride=c["stride"], allpairs=allpairs, return_feat=True, **extra)
    assert losses is None
    assert len(preds) == c["iters"] and len(preds2) == c["iters"] + 4
    got = torch.stack(preds).numpy()
    ref = GOLD[name + "/preds"]
    assert got.shape == ref.shape
    tol = TOL_PX if c["head_scale"] < 1.0 else 5e-3
    assert np.abs(got - ref).max() < tol, np.abs(got - ref).max()
    assert np.abs(ffeat.numpy() - GOLD[name + "/ffeat"]).max() < 1e-5
    assert np.abs(vis_e.numpy() - GOLD[name + "/vis_e"]).max() < 2e-3
    # anim list layout, nets/pips.py:474-475, :562-563
    assert torch.equal(preds2[0], preds2[1]) and torch.equal(preds2[-1], preds2[-2]) and torch.equal(preds2[-1], preds[-1])


def test_state_dict_spec_counts():
    spec = po.state_dict_spec()
    assert len(spec) == 200
    total = sum(int(np.prod(s)) for _, s in spec)
    assert total == 28677713


def test_corr_local_equals_allpairs():
    torch.manual_seed(0)
    B, S, N, C, H, W = 2, 8, 9, 128, 20, 28
    fmaps = torch.randn(B, S, C, H, W)
    pyr = po.build_pyramid(fmaps)
    coords = torch.rand(B, S, N, 2) * torch.tensor([W + 8.0, H + 8.0]) - 4.0
    coords[0, 0, 0] = torch.tensor([3.0, 5.0])
    tg = torch.randn(B, S, N, C)
    a = po.sample_allpairs(po.corr_allpairs(pyr, tg), coords)
    b = po.corr_local(pyr, tg, coords)
    assert a.shape == (B, S, N, 196)
    assert (a - b).abs().max() < 2e-4


def test_times_axis_is_linspace_0_S():
    t = po.times_axis(8)
    assert t[0] == 0 and t[-1] == 8 and abs(float(t[1]) - 8 / 7) < 1e-6


def test_oracle_score_maps_match_reference():
    """The dense score map fcps (nets/pips.py:504-511) of the oracle against the reference's own, recorded for
    three particles of the supervised case (captured as the argument of score_map_loss, nets/pips.py:603)."""
    from tests.golden.make_golden import FCP_SEL, LOSS_CASE
    c = LOSS_CASE
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, extra = case_inputs(c)
    traces = []
    with torch.no_grad():
        po.forward(sd, xys, rgbs, iters=c["iters"], stride=c["stride"], traces=traces, **extra)
        x = 2 * (rgbs / 255.0) - 1.0
        B, S = rgbs.shape[:2]
        H8, W8 = c["H"] // c["stride"], c["W"] // c["stride"]
        fmaps = po.fnet(sd, x.reshape(B * S, 3, c["H"], c["W"]), c["stride"]).reshape(B, S, po.LATENT, H8, W8)
        pyr = po.build_pyramid(fmaps)
        got = torch.stack([po.heatmap_fcp(po.corr_allpairs(pyr, tr["ffeats_in"]), H8, W8)[:, :, FCP_SEL] for tr in traces], dim=2)
    ref = GOLD["loss_s8/fcps_sel"]
    assert got.shape == ref.shape
    err = np.abs(got.numpy() - ref).max()
    print("oracle fcps vs reference: max err", err, "|fcps| max", np.abs(ref).max())
    assert err < 2e-4


def test_oracle_matches_reference_at_bench_size():
    """BASELINE cfg 2 at full size (B=4, 8x384x512, N=1024, 6 iterations): the oracle's local-correlation formulation
    against the reference recorded once by make_golden.py --cfg2.  About a minute of CPU time."""
    from tests.golden.make_golden import CFG2_CASE as c, CFG2_EVERY
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_cfg2.npz"))
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, _ = case_inputs(c)
    with torch.no_grad():
        preds, _, vis_e, ffeat, _ = po.forward(sd, xys, rgbs, iters=c["iters"], stride=c["stride"], return_feat=True)
    p = torch.stack(preds).numpy()
    err = max(np.abs(p[:, :, :, ::CFG2_EVERY] - gold["preds_sub"]).max(), np.abs(p[-1] - gold["preds_final"]).max())
    print("oracle vs reference at cfg2: max err px", err)
    assert err < TOL_PX
    assert np.abs(vis_e.numpy() - gold["vis_e"]).max() < 2e-3
    assert np.abs(ffeat[:, ::16].numpy() - gold["ffeat_sub"]).max() < 1e-5


def test_oracle_matches_reference_on_the_real_demo_clip():
    """BASELINE cfg 1 on the real frames demo_images/000100-000107.jpg (stored as JPEG bytes inside the fixture), the
    recipe of demo.py:21-41 (16 x 16 grid, 360 x 640, stride 4, 6 iterations): oracle against the reference's recording."""
    import hashlib
    from tests.golden.make_golden import DEMO_CASE as c, demo_decode, demo_inputs
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_demo.npz"))
    raw = demo_decode([gold[f"jpeg{i}"].tobytes() for i in range(8)])
    if hashlib.sha256(raw.to(torch.uint8).numpy().tobytes()).hexdigest() != str(gold["pixels_sha256"]):
        pytest.skip("this host's JPEG decoder produces different pixels than the one the fixture was recorded with")
    rgbs, xy = demo_inputs(raw)
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    with torch.no_grad():
        preds, _, vis_e, ffeat, _ = po.forward(sd, xy, rgbs, iters=c["iters"], stride=c["stride"], return_feat=True)
    err = np.abs(torch.stack(preds).numpy() - gold["preds"]).max()
    print("oracle vs reference on the demo clip: max err px", err)
    assert err < TOL_PX
    assert np.abs(vis_e.numpy() - gold["vis_e"]).max() < 2e-3
    assert np.abs(ffeat.numpy() - gold["ffeat"]).max() < 1e-5


def test_oracle_matches_reference_at_cfg4_shape():
    """BASELINE cfg 4 shape (8 x 720 x 1280, stride 8 -> 90 x 160 maps): the 256-particle chunk the reference was run on
    (every 64th of the 16384 queries; test_on_davis.py:111-125 chunks the same way)."""
    from tests.golden.make_golden import CFG4_CASE as c, CFG4_EVERY
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_cfg4.npz"))
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, _ = case_inputs(c)
    with torch.no_grad():
        preds, _, vis_e, ffeat, _ = po.forward(sd, xys[:, ::CFG4_EVERY].contiguous(), rgbs, iters=c["iters"], stride=c["stride"],
                                               return_feat=True)
    err = np.abs(torch.stack(preds).numpy() - gold["preds"]).max()
    print("oracle vs reference at the cfg4 shape: max err px", err)
    assert err < TOL_PX
    assert np.abs(vis_e.numpy() - gold["vis_e"]).max() < 2e-3
    assert np.abs(ffeat.numpy() - gold["ffeat"]).max() < 1e-5
