"""BASELINE cfg5: chained tracking over a 100-frame 360x640 clip, N=512, 8-frame windows, stride 4, 1 GPU."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import synthetic
from pips_b200.chain import track_chain

dev = torch.device("cuda", 0)
T, H, W, N = 100, 360, 640, 512
rgbs = synthetic.smooth_video(1, T, H, W, seed=99).to(dev)
xy0 = synthetic.random_queries(1, N, H, W, seed=98).to(dev)
model = synthetic.seeded_model(stride=4).to(dev).eval()
for _ in range(2):
    trajs, rounds = track_chain(model, rgbs, xy0, iters=6, return_rounds=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
trajs, rounds = track_chain(model, rgbs, xy0, iters=6, return_rounds=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"chain cfg5: T={T} {H}x{W} N={N} stride 4: {rounds} batched rounds, {dt*1e3:.1f} ms wall, "
      f"{T*N/dt:.0f} tracked particle-frames/s; finite={bool(torch.isfinite(trajs).all())}")

# data-independent schedule (fixed advance 7: ceil((T-1)/7) rounds) -- the throughput figure of SURVEY.md section 8d
for _ in range(2):
    track_chain(model, rgbs, xy0, iters=6, advance=7)
torch.cuda.synchronize()
t0 = time.perf_counter()
trajs, rounds = track_chain(model, rgbs, xy0, iters=6, return_rounds=True, advance=7)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"chain cfg5, fixed advance 7: {rounds} rounds, {dt*1e3:.1f} ms wall, {T*N/dt:.0f} tracked particle-frames/s")
