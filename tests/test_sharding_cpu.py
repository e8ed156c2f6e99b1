"""Host logic of the particle-sharded path with world_size 2 on CPU (gloo): shard bounds, tail
padding, per-iteration all-gather and reassembly.  The CUDA engine is replaced by a deterministic
stand-in (the kernels themselves are covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pips_b200.sharding import encode_sharded, gather_mode, refine_sharded, shard_bounds


def test_shard_bounds_cover_all_particles():
    for N in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                n0, n1, per = shard_bounds(N, r, world)
                assert 0 <= n0 <= n1 <= N and n1 - n0 <= per
                seen += list(range(n0, n1))
            assert seen == list(range(N))


class _FakeEngine:
    def refine(self, module, fmaps, coords, feat_init, iters, stride, on_iter=None):
        B, S, n, _ = coords.shape
        preds = torch.stack([coords * stride + (it + 1) for it in range(iters)])
        for it in range(iters):
            if on_iter is not None:
                on_iter(it, preds[it].contiguous())
        vis = coords.sum(-1)
        ffeat = coords[:, 0, :, :1].repeat(1, 1, 128) if feat_init is None else feat_init * 2
        return preds, vis, ffeat


class _FakeModel:
    def __init__(self, shard):
        self._shard = shard
        self.engine = _FakeEngine()

    def encode(self, rgbs):                      # per-frame function of the input, channels-last like the real encoder
        B, S, C, H, W = rgbs.shape
        f = rgbs.reshape(B * S, C, H, W).mean(1, keepdim=True).repeat(1, 128, 1, 1)[:, :, ::8, ::8]
        return f.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).reshape(B, S, 128, H // 8, W // 8)


def _worker(rank, world, port, N, use_feat, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B, S, iters = 2, 8, 3
        coords = torch.randn(B, S, N, 2)
        feat = torch.randn(B, N, 128) if use_feat else None
        model = _FakeModel((rank, world, None))
        preds, vis, ff = refine_sharded(model, torch.zeros(B, S, 128, 4, 4), coords, feat, iters, 4.0)
        exp_p, exp_v, exp_f = _FakeEngine().refine(None, None, coords, feat, iters, 4.0)
        ok = torch.equal(preds, exp_p) and torch.equal(vis, exp_v) and torch.equal(ff, exp_f)
        rgbs = torch.rand(B, 3, 3, 16, 24)       # 6 frames over 2 ranks; also an odd count: 3 frames
        for clip in (rgbs, rgbs[:1]):
            ok = ok and torch.equal(encode_sharded(model, clip), model.encode(clip))
        # exchange mode: both ranks on this host -> peer slabs (taken for CUDA tensors only; the CPU tensors above went
        # through the all-gather path); the environment override wins and is cached per model
        ok = ok and gather_mode(model) == "p2p"
        os.environ["PIPS_B200_GATHER"] = "nccl"
        ok = ok and gather_mode(_FakeModel((rank, world, None))) == "nccl" and gather_mode(model) == "p2p"
        del os.environ["PIPS_B200_GATHER"]
        q.put((rank, bool(ok), tuple(preds.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,use_feat", [(10, False), (7, True), (1, False)])
def test_refine_sharded_world2_gloo(N, use_feat):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, use_feat, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (3, 2, 8, N, 2) for _, _, shape in res)


def test_speed_weighted_shard_sizes():
    """sharding.shard_sizes: deterministic, sums to N, every rank gets work, proportional to the measured rates, whole
    32-track GEMM tiles where N allows."""
    from pips_b200.sharding import shard_sizes
    rates = [1.0, 1.1, 0.9, 1.0, 1.05, 0.95, 1.0, 1.0]
    sz = shard_sizes(8192, rates)
    assert sum(sz) == 8192 and all(s % 32 == 0 for s in sz) and sz == shard_sizes(8192, list(rates))
    assert sz[1] > sz[0] > sz[2] and abs(sz[1] / 8192 - 1.1 / 8.0) < 0.01
    assert shard_sizes(17, [1, 1, 1, 1]) == [5, 4, 4, 4]                  # small N: single tracks, still sums to N
    assert sum(shard_sizes(20, [1.0, 1.2])) == 20 and shard_sizes(20, [1.0, 1.2])[1] > 10
    assert shard_sizes(3, [1, 1, 1, 1]) is None                           # fewer particles than ranks: equal padded shards
    assert sum(shard_sizes(4097, [3.0, 1.0])) == 4097 and min(shard_sizes(4097, [30.0, 1.0])) >= 1
