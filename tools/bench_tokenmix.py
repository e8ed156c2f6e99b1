"""Micro-benchmark of pips_tokenmix (and the tc variant) at the bench size: 4096 tracks."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import _lib as L
lib = L.load()
dev = "cuda:0"
seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
x = torch.randn(seqs * 8, 512, device=dev)
p = {k: torch.randn(n, device=dev) * 0.2 + (1.0 if k.endswith("w") and k.startswith("ln") else 0.0)
     for k, n in dict(ln1w=512, ln1b=512, w1=256, b1=32, w2=256, b2=8, ln2w=512, ln2b=512).items()}
yh = torch.empty(seqs * 8, 512, dtype=torch.bfloat16, device=dev); yl = torch.empty_like(yh)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(fn, name):
    args = (L.ptr(x), seqs, L.ptr(p["ln1w"]), L.ptr(p["ln1b"]), L.ptr(p["w1"]), L.ptr(p["b1"]), L.ptr(p["w2"]), L.ptr(p["b2"]),
            L.ptr(p["ln2w"]), L.ptr(p["ln2b"]), L.ptr(yh), L.ptr(yl))
    for _ in range(3): L.check(fn(*args))
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.check(fn(*args)); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); print(f"{name}: median {ts[10]*1e3:.1f} us  min {ts[0]*1e3:.1f} us")
# pips_tokenmix dispatches on PIPS_B200_TOKENMIX (read once per process): "simt" = CUDA-core kernel, default = tensor-core kernel
run(lambda *a: lib.pips_tokenmix(*a, None, st), f"pips_tokenmix [PIPS_B200_TOKENMIX={os.environ.get('PIPS_B200_TOKENMIX', 'tc (default)')}]")
run(lambda *a: lib.pips_tokenmix_tc(*a, st), "pips_tokenmix_tc (tensor-core kernel, explicit entry)")
