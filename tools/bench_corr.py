"""Micro-benchmark of pips_corr_gather at the bench size (B=4, S=8, N=1024, 48x64 maps -> 32768 units) and, with
--big, at a pyramid larger than L2 (BASELINE cfg 5: 100 frames of 90x160 maps at stride 4, fp32: 0.98 GB)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import _lib as L
from pips_b200.engine import Pyramid

lib = L.load()
dev = torch.device("cuda", 0)
big = "--big" in sys.argv
feat = L.FEAT_BF16 if "--bf16" in sys.argv else L.FEAT_F32
torch.manual_seed(0)
if big:
    B, T, N, H8, W8 = 1, 100, 4096, 90, 160          # chained windows: every track reads its own 8 frames of the clip
else:
    B, T, N, H8, W8 = 4, 8, 1024, 48, 64
S = 8
fmaps = torch.randn(B * T, 128, H8, W8, device=dev)
pyr = Pyramid(B * T, H8, W8, feat, dev)
st = torch.cuda.current_stream().cuda_stream
pyr.build(fmaps, st)
coords = torch.rand(B, S, N, 2, device=dev) * torch.tensor([W8 - 1.0, H8 - 1.0], device=dev)
ffeats = torch.randn(B * N, S, 128, device=dev)
times = torch.linspace(0, S, S, device=dev)
fb = torch.randint(0, T - 8, (B, N), device=dev, dtype=torch.int32) if big else None
M = B * N * S
x_hi = torch.empty(M, 576, dtype=torch.bfloat16, device=dev); x_lo = torch.empty_like(x_hi)
lvl = L.ptr_array(pyr.levels())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    L.check(lib.pips_corr_gather(lvl, feat, B, S, N, H8, W8, L.ptr(coords), L.ptr(ffeats), L.ptr(times), L.ptr(fb), T if big else 0,
                                 L.ptr(x_hi), L.ptr(x_lo), None, 576, st))
for _ in range(3): run()
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
e_f = 4 if feat == L.FEAT_F32 else 2
pyr_bytes = sum(t.numel() * t.element_size() for t in pyr.levels())
b_f32 = 4 * 64 * 128 * e_f + 128 * 4 + 576 * 4
b_surv = 66440
print(f"corr_gather {'big (pyramid %.0f MB > L2)' % (pyr_bytes / 1e6) if big else 'cfg2 (pyramid %.0f MB, L2-resident)' % (pyr_bytes / 1e6)} "
      f"feat={'fp32' if feat == L.FEAT_F32 else 'bf16'} units={M}: median {ts[10]*1e3:.1f} us  min {ts[0]*1e3:.1f} us; "
      f"algorithmic {b_f32} B/unit -> {b_f32 * M / ts[10] / 1e6:.0f} GB/s; SURVEY 8d count {b_surv} B/unit -> {b_surv * M / ts[10] / 1e6:.0f} GB/s")
