"""torchrun --nproc-per-node G tools/check_sharded.py : particle-sharded forward == single-GPU forward."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
CASES = {"rect_s4": dict(B=2, H=128, W=192, N=20, stride=4, iters=6, warm=False),
         "warm_s8": dict(B=1, H=128, W=160, N=10, stride=8, iters=2, warm=True),
         "odd_s8": dict(B=1, H=184, W=360, N=17, stride=8, iters=4, warm=False)}
for name, c in CASES.items():
    rgbs = synthetic.smooth_video(c["B"], 8, c["H"], c["W"], seed=5)
    xys = synthetic.random_queries(c["B"], c["N"], c["H"], c["W"], seed=6)
    extra = {}
    if c["warm"]:
        g = torch.Generator().manual_seed(7)
        extra = {"coords_init": (xys[:, None] + torch.cumsum(torch.randn(c["B"], 8, c["N"], 2, generator=g), 1)).to(dev),
                 "feat_init": (torch.randn(c["B"], c["N"], 128, generator=g) * 0.5).to(dev)}
    single = synthetic.seeded_model(stride=c["stride"], seed=3).to(dev).eval()
    with torch.no_grad():
        a = single(xys.to(dev), rgbs.to(dev), iters=c["iters"], return_feat=True, **extra)
    for mode in ("p2p", "nccl"):            # peer-mapped slabs (stores fused into the update kernel) / NCCL all-gathers
        sharded = synthetic.seeded_model(stride=c["stride"], seed=3).to(dev).eval()
        sharded.shard_particles(balance=True)     # speed-weighted shares after the 5th call: uneven shards must be bit-exact too
        sharded._gather_mode = mode
        for rep in range(8):                # repeated calls reuse the slab (the barriers must fence it); after the 5th the
                                            # shards are speed-weighted (sharding._Balance): uneven sizes, still bit-exact
            with torch.no_grad():
                b = sharded(xys.to(dev), rgbs.to(dev), iters=c["iters"], return_feat=True, **extra)
            same = all(torch.equal(x, y) for x, y in zip(a[0], b[0])) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
            err = max((x - y).abs().max().item() for x, y in zip(a[0], b[0]))
            ok = ok and same
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        t0.record()
        for rep in range(5):
            with torch.no_grad():
                sharded(xys.to(dev), rgbs.to(dev), iters=c["iters"], return_feat=True, **extra)
        t1.record()
        torch.cuda.synchronize()
        print(f"rank {rank}/{world} {name} [{mode}]: N={c['N']} sharded==single bit-exact: {same} (max diff {err:.2e}); "
              f"{t0.elapsed_time(t1) / 5:.2f} ms/forward", flush=True)
        sharded.close_peer_slabs()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
