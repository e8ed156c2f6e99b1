"""One forward of the demo shape (BASELINE cfg1: B=1, 8x360x640, N=256, stride 4) inside cudaProfilerStart/Stop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rgbs = synthetic.smooth_video(1, 8, 360, 640, seed=3).to(torch.bfloat16).to(dev)
xys = synthetic.random_queries(1, N, 360, 640, seed=4).to(dev)
model = synthetic.seeded_model(stride=4).to(dev).eval()
with torch.no_grad():
    for _ in range(2):
        model(xys, rgbs, iters=6)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(xys, rgbs, iters=6)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one demo-shape forward, N =", N)
