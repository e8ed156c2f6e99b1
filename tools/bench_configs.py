"""Other BASELINE.json configurations on one GPU: cfg1 (demo shape), cfg4's per-rank share, sanity + wall time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import synthetic

dev = torch.device("cuda", 0)
CFG = {
    "cfg1 demo shape (B=1, 8x360x640, N=256, stride 4)": dict(B=1, H=360, W=640, N=256, stride=4),
    "cfg1 demo shape, stride 8": dict(B=1, H=360, W=640, N=256, stride=8),
    "cfg4 per-rank share of 8 (B=1, 8x720x1280, N=2048, stride 8)": dict(B=1, H=720, W=1280, N=2048, stride=8),
    "cfg4 whole on one GPU (N=16384)": dict(B=1, H=720, W=1280, N=16384, stride=8),
}
for name, c in CFG.items():
    rgbs = synthetic.smooth_video(c["B"], 8, c["H"], c["W"], seed=3).to(torch.bfloat16).to(dev)
    xys = synthetic.random_queries(c["B"], c["N"], c["H"], c["W"], seed=4).to(dev)
    model = synthetic.seeded_model(stride=c["stride"]).to(dev).eval()
    with torch.no_grad():
        for _ in range(3):
            out = model(xys, rgbs, iters=6)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            out = model(xys, rgbs, iters=6)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
    upd = c["B"] * 8 * c["N"] * 6
    print(f"{name}: {dt*1e3:8.2f} ms/forward  {upd/dt/1e6:6.2f} M updates/s  finite={bool(torch.isfinite(out[0][-1]).all())} "
          f"mean|d|={float((out[0][-1]-xys[:,None]).abs().mean()):.3f}px  mem={torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    del model
    torch.cuda.empty_cache()
