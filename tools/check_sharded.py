"""torchrun --nproc-per-node G tools/check_sharded.py : particle-sharded forward == single-GPU forward."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pips_oracle as po          # input generator / weights only
from pips_b200 import Pips
from tests.golden.make_golden import CASES, case_inputs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for name in ("rect_s4_oob", "warm_s8", "odd_s8"):
    c = CASES[name]
    sd = po.init_state_dict(seed=c["seed"], head_scale=c["head_scale"])
    rgbs, xys, extra = case_inputs(c)
    extra = {k: v.to(dev) for k, v in extra.items()}
    single = Pips(S=8, stride=c["stride"]).to(dev).eval()
    single.load_state_dict(sd)
    sharded = Pips(S=8, stride=c["stride"]).to(dev).eval()
    sharded.load_state_dict(sd)
    sharded.shard_particles()
    with torch.no_grad():
        a = single(xys.to(dev), rgbs.to(dev), iters=c["iters"], return_feat=True, **extra)
        b = sharded(xys.to(dev), rgbs.to(dev), iters=c["iters"], return_feat=True, **extra)
    same = all(torch.equal(x, y) for x, y in zip(a[0], b[0])) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    err = max((x - y).abs().max().item() for x, y in zip(a[0], b[0]))
    print(f"rank {rank}/{world} {name}: N={c['N']} sharded==single bit-exact: {same} (max diff {err:.2e})", flush=True)
    ok = ok and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
