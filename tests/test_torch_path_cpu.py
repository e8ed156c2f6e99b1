"""The supervised / training path of Pips.forward (pips_b200/torch_path.py) against the reference's
recorded outputs, on CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import pips_oracle as po
from pips_b200 import Pips
from tests.golden.make_golden import LOSS_CASE, case_inputs, loss_targets

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.npz"))


@pytest.mark.parametrize("name,is_train", [("loss_s8", False), ("train_s8", True)])
def test_supervised_forward_matches_reference(name, is_train):
    c = LOSS_CASE
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    model = Pips(S=8, stride=c["stride"]).eval()
    model.load_state_dict(sd, strict=True)
    rgbs, xys, _ = case_inputs(c)
    trajs_g, vis_g, valids = loss_targets(c, xys)
    with torch.no_grad():
        preds, preds2, vis_e, losses = model(xys, rgbs, iters=c["iters"], trajs_g=trajs_g, vis_g=vis_g, valids=valids, is_train=is_train)
    assert len(preds2) == c["iters"] + 4
    assert np.abs(torch.stack(preds).numpy() - GOLD[name + "/preds"]).max() < 1e-3
    assert np.abs(vis_e.numpy() - GOLD[name + "/vis_e"]).max() < 2e-3
    got = np.array([float(l) for l in losses])
    assert np.allclose(got, GOLD[name + "/losses"], rtol=1e-4, atol=1e-5), (got, GOLD[name + "/losses"])


def test_training_path_has_gradients():
    c = LOSS_CASE
    model = Pips(S=8, stride=8).train()
    rgbs, xys, _ = case_inputs(c)
    trajs_g, vis_g, valids = loss_targets(c, xys)
    preds, _, vis_e, losses = model(xys[:1, :3], rgbs[:1], iters=1, trajs_g=trajs_g[:1, :, :3], vis_g=vis_g[:1, :, :3],
                                    valids=valids[:1, :, :3], is_train=True)
    sum(losses).backward()
    assert model.delta_block.to_delta[0].weight.grad is not None
    assert model.fnet.conv1.weight.grad is not None
