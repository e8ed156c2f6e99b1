"""One warm-up forward + one forward inside cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
Same workload as bench.py at N=1 (BASELINE cfg2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pips_b200 import synthetic  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
feat = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda", 0)
rgbs, xys = bench.make_inputs(bench.N_PER_GPU)
model = synthetic.seeded_model(stride=bench.STRIDE, precision=precision, feat_dtype=feat).to(dev).eval()
rgbs, xys = rgbs.to(dev), xys.to(dev)
with torch.no_grad():
    model(xys, rgbs, iters=bench.ITERS)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(xys, rgbs, iters=bench.ITERS)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward:", precision, feat, "launches", model.engine.launches)
