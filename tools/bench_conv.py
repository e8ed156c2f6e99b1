"""Micro-benchmark of the layer-1 convolution (64 -> 64, 3x3, 32 x 192 x 256 maps as at BASELINE cfg 2): the row-ring
kernel (csrc/conv_rows.cu, with InstanceNorm partial statistics) against the tap-by-tap kernel (csrc/conv_tc.cu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200.encoder_fast import _Pair, conv_rows, conv_tc
dev = "cuda:0"
N, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 192, 256)
torch.manual_seed(0)
conv = torch.nn.Conv2d(64, 64, 3, padding=1).to(dev)
x = torch.randn(N, H, W, 64, device=dev)
pair = _Pair(N, H, W, 64, dev); pair.hi.copy_(x.to(torch.bfloat16)); pair.lo.copy_((x - pair.hi.float()).to(torch.bfloat16))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(fn, name):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    fl = 2.0 * N * H * W * 64 * 64 * 9
    print(f"{name}: median {ts[5]*1e3:.1f} us  min {ts[0]*1e3:.1f} us  ({fl / ts[5] / 1e9:.0f} TFLOP/s algorithmic, x3 issued)")
run(lambda: conv_rows(pair, conv), f"conv_rows {N}x{H}x{W} (+ statistics finalize)")
run(lambda: conv_tc(pair, conv), f"conv_tc   {N}x{H}x{W}")
