"""Headless equivalent of the reference's demo.py:21-41 (run_model) with the drop-in module.

demo.py itself needs imageio / tensorboardX / cv2 and the Dropbox checkpoint; this driver performs exactly its
model-facing steps -- resize the 8-frame clip to 360x640, lay out a 16x16 query grid, call
``model(xy, rgbs, iters=6)``, take ``preds[-1]`` -- on a synthetic clip (or on .npy frames given on the command
line), with ``from pips_b200 import Pips`` in place of ``from nets.pips import Pips``.

    python examples/demo_headless.py [frames.npy] [checkpoint_dir]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import Pips, synthetic          # noqa: E402   (reference: from nets.pips import Pips)


def run_model(model, rgbs, N):
    rgbs = rgbs.cuda().float()                                     # B, S, C, H, W          demo.py:22
    B, S, C, H, W = rgbs.shape
    rgbs_ = F.interpolate(rgbs.reshape(B * S, C, H, W), (360, 640), mode="bilinear")      # demo.py:26-29
    H, W = 360, 640
    rgbs = rgbs_.reshape(B, S, C, H, W)
    N_ = int(np.sqrt(N).round())                                   # uniform grid           demo.py:32-36
    gy, gx = torch.meshgrid(torch.arange(N_, device="cuda").float(), torch.arange(N_, device="cuda").float(), indexing="ij")
    gy = 8 + gy.reshape(1, -1) / float(N_ - 1) * (H - 16)
    gx = 8 + gx.reshape(1, -1) / float(N_ - 1) * (W - 16)
    xy = torch.stack([gx, gy], dim=-1).repeat(B, 1, 1)             # B, N, 2
    preds, preds_anim, vis_e, stats = model(xy, rgbs, iters=6)     #                        demo.py:40
    return preds[-1], vis_e, preds_anim


def main():
    if len(sys.argv) > 1:
        rgbs = torch.from_numpy(np.load(sys.argv[1])).permute(0, 3, 1, 2).unsqueeze(0)    # (S,H,W,3) uint8 -> 1,S,3,H,W
    else:
        rgbs = synthetic.smooth_video(1, 8, 480, 854, seed=0)
    model = Pips(stride=4).cuda()                                  #                        demo.py:114
    if len(sys.argv) > 2:
        sys.path.insert(0, "/root/reference")
        import saverloader                                         # the reference's loader works unchanged
        saverloader.load(sys.argv[2], model)
    else:
        model = synthetic.seeded_model(stride=4).cuda()
    model.eval()
    with torch.no_grad():
        trajs_e, vis_e, anim = run_model(model, rgbs, N=16 ** 2)
    torch.cuda.synchronize()
    print("trajs_e", tuple(trajs_e.shape), "min %.2f max %.2f" % (float(trajs_e.min()), float(trajs_e.max())),
          "| vis_e", tuple(vis_e.shape), "| animation frames", len(anim))


if __name__ == "__main__":
    main()
