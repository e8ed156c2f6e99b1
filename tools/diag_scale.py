"""torchrun --nproc-per-node G tools/diag_scale.py : where a particle-sharded bench step spends its time, per rank.

Per phase of the sharded forward (pips_b200/sharding.py marks), GPU-time medians over the steps of a run-ahead loop
(no host sync between steps, like bench.py's device-timed region) and of a loop with a host sync per step, for the
peer-slab and the NCCL result exchange; plus each rank's SM clock / power under load (NVML)."""
import os
import statistics
import sys
import threading
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pips_b200 import sharding, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S, H, W, NPER, ITERS, STEPS = 4, 8, 384, 512, 1024, 6, 25
rgbs = synthetic.smooth_video(B, S, H, W, seed=1234).to(torch.bfloat16).to(dev)
xys = synthetic.random_queries(B, NPER * world, H, W, seed=4321).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


class Nvml(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.mhz, self.watts, self.stop = [], [], False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            while not self.stop:
                self.mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                self.watts.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                time.sleep(0.05)
        except Exception as e:                                       # noqa: BLE001
            self.mhz.append(-1)


def loop(model, sync_each):
    sharding.TRACE = []
    marks = []
    t0 = time.perf_counter()
    for _ in range(STEPS):
        flush.zero_()
        n0 = len(sharding.TRACE)
        with torch.no_grad():
            model(xys, rgbs, iters=ITERS)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        sharding.TRACE.append(("end", ev))
        marks.append((n0, len(sharding.TRACE)))
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / STEPS * 1e3
    tr, sharding.TRACE = sharding.TRACE, None
    phases = {}
    for a, b in marks:
        seg = tr[a:b]
        for (l0, e0), (l1, e1) in zip(seg[:-1], seg[1:]):
            phases.setdefault(l1, []).append(e0.elapsed_time(e1))
        phases.setdefault("step", []).append(seg[0][1].elapsed_time(seg[-1][1]))
    return {k: statistics.median(v) for k, v in phases.items()}, wall


for mode in ("p2p", "nccl"):
    model = synthetic.seeded_model(stride=8, seed=0).to(dev).eval()
    model.shard_particles(balance=os.environ.get("PIPS_B200_BALANCE", "0") == "1")
    model._gather_mode = mode
    for _ in range(8):                                   # > the 5 calls after which the shards become speed-weighted
        with torch.no_grad():
            model(xys, rgbs, iters=ITERS)
    bal = getattr(model, "_balance", None)
    if rank == 0:
        print(f"=== {mode}: shard sizes", sharding.shard_sizes(NPER * world, bal.weights) if bal is not None and bal.weights else "equal",
              "rates (particles/ms)", [round(w, 1) for w in bal.weights] if bal is not None and bal.weights else None, flush=True)
    for sync_each in (False, True):
        dist.barrier()
        torch.cuda.synchronize()
        mon = Nvml()
        mon.start()
        ph, wall = loop(model, sync_each)
        mon.stop = True
        mon.join(timeout=1)
        row = {"rank": rank, "mode": mode, "sync_each": sync_each, "wall_ms": round(wall, 2),
               "mhz": statistics.median(mon.mhz) if mon.mhz else None, "watts": round(statistics.median(mon.watts)) if mon.watts else None,
               **{k: round(v, 2) for k, v in ph.items()}}
        rows = [None] * world
        dist.all_gather_object(rows, row)
        if rank == 0:
            print(f"--- {mode}, {'host sync per step' if sync_each else 'run-ahead'} ---")
            for r in rows:
                print(r, flush=True)
    model.close_peer_slabs()
dist.barrier()
dist.destroy_process_group()
