"""fnet variants on the bench workload: time and error against strict-fp32 cuDNN."""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pips_b200 import Pips
from pips_b200 import encoder as enc

dev = torch.device("cuda", 0)
sd, rgbs, xys = bench.make_inputs(bench.N_PER_GPU)
model = Pips(S=8, stride=8).to(dev).eval()
model.load_state_dict(sd)
x = (2 * (rgbs.to(dev).float() / 255.0) - 1.0).reshape(32, 3, 384, 512)

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

def run(mode, tf32, bench_flag, cl=False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = bench_flag
    model.fnet.mode = mode
    xx = x.contiguous(memory_format=torch.channels_last) if cl else x
    if cl: model.fnet.to(memory_format=torch.channels_last)
    else: model.fnet.to(memory_format=torch.contiguous_format)
    with torch.no_grad():
        return timeit(lambda: model.fnet(xx))

t0, ref = run("plain", False, False)
print(f"strict fp32             {t0:8.2f} ms   |fmaps| max {ref.abs().max().item():.3f} mean {ref.abs().mean().item():.3f}")
for name, args in [("fp32 + cudnn.benchmark", ("plain", False, True)), ("tf32 plain", ("plain", True, False)),
                   ("tf32 plain + benchmark", ("plain", True, True)), ("tf32 plain CL + bench", ("plain", True, True, True)),
                   ("3xTF32 NCHW", ("x3", True, False)), ("3xTF32 NCHW + bench", ("x3", True, True)),
                   ("3xTF32 CL + bench", ("x3", True, True, True))]:
    try:
        t, out = run(*args)
        print(f"{name:24s}{t:8.2f} ms   max|err| {(out.float() - ref).abs().max().item():.3e}  mean|err| {(out.float() - ref).abs().mean().item():.3e}")
    except Exception as e:
        print(name, "FAILED", repr(e)[:300])
