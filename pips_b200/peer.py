"""Peer-mapped result slabs for particle-sharded runs on one node.

Each rank owns one device allocation (``pips_peer_alloc``: cudaMalloc + CUDA IPC handle) that every other rank
maps (``pips_peer_open``).  The update kernel stores its slice of each iteration's prediction straight into all
slabs over NVLink (``pips_peer_out`` of include/pips_b200.h), ``pips_peer_scatter`` does the same for the
visibility logits and the carried features, and ``pips_peer_barrier`` -- a flag barrier in the slabs themselves --
replaces the synchronisation a collective would imply.  torch.distributed is used once, to exchange the 64-byte
handles; nothing on the per-forward path calls a collective library.

Slab layout (fp32 words): [0, 64) barrier flags (int32, one per rank) | coords (iters,B,S,n_total,2) |
vis (B,S,n_total) | ffeat (B,n_total,128).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib as L

FLAG_WORDS = 64
LATENT = 128


class _DevPtr:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface, version 2)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerSlab:
    _generation = 0          # distinguishes slabs that happen to get the same device address from cudaMalloc

    def __init__(self, nbytes: int, rank: int, world: int, group, device: torch.device):
        lib = L.load()
        PeerSlab._generation += 1
        self.generation = PeerSlab._generation
        if world > L.MAX_PEERS:
            raise L.PipsCudaError(f"pips_b200: peer slabs support up to {L.MAX_PEERS} ranks per node")
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.nbytes = (nbytes + 255) // 256 * 256
        self.epoch = 0
        self.timeout_ms = int(os.environ.get("PIPS_B200_PEER_TIMEOUT_MS", "120000"))
        # Every step that can fail locally is followed by an exchange of its outcome, so that all ranks either
        # finish the set-up or raise together (a one-sided failure would leave the others in a collective).
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self.local, self.ptrs = 0, []
        err = None
        with torch.cuda.device(device):
            if lib.pips_peer_alloc(self.nbytes, C.byref(ptr), handle) != 0:
                err = lib.pips_last_error().decode()
            else:
                self.local = ptr.value
            handles: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(handles, None if err else handle.raw, group=group)
            if all(h is not None for h in handles):
                for r in range(world):
                    if r == rank:
                        self.ptrs.append(self.local)
                        continue
                    p = C.c_void_p()
                    if lib.pips_peer_open(handles[r], C.byref(p)) != 0:
                        err = lib.pips_last_error().decode()
                        break
                    self.ptrs.append(p.value)
            elif err is None:
                err = "another rank could not allocate its slab"
            errs: List[Optional[str]] = [None] * world
            dist.all_gather_object(errs, err, group=group)
            if any(e is not None for e in errs):
                for r, p in enumerate(self.ptrs):
                    if r != rank:
                        lib.pips_peer_close(p)
                if self.local:
                    lib.pips_peer_free(self.local)
                raise L.PipsCudaError("pips_b200: peer slab set-up failed: " +
                                      "; ".join(f"rank {r}: {e}" for r, e in enumerate(errs) if e))
        self._keep = _DevPtr(self.local, self.nbytes // 4)
        self.words = torch.as_tensor(self._keep, device=device)            # local slab as a float32 vector
        self._flag_ptrs = (C.c_void_p * world)(*self.ptrs)
        dist.barrier(group=group)                                           # every mapping exists before first use

    def region_ptrs(self, word_offset: int):
        """ctypes array with the address of ``word_offset`` inside every rank's slab."""
        return (C.c_void_p * self.world)(*[p + 4 * word_offset for p in self.ptrs])

    def barrier(self) -> None:
        """Enqueue the flag barrier on the current stream (all ranks must call it the same number of times)."""
        self.epoch += 1
        L.check(L.load().pips_peer_barrier(self._flag_ptrs, self.rank, self.world, self.epoch, self.timeout_ms,
                                           torch.cuda.current_stream(self.device).cuda_stream), "pips_peer_barrier")

    def close(self) -> None:
        """Collective: unmap the peers' slabs, then free the local one."""
        lib = L.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                L.check(lib.pips_peer_close(p), "pips_peer_close")
        dist.barrier(group=self.group)
        self.words = None
        L.check(lib.pips_peer_free(self.local), "pips_peer_free")
        self.ptrs, self.local = [], 0


class PeerPlan:
    """Where one forward's results go inside the slabs: handed to RefineEngine.refine(peer=...).
    Equal shards of ``per`` particles (the tail padded), or -- ``sizes`` given -- one shard of ``sizes[r]`` particles
    per rank laid out back to back (speed-weighted sharding, no padding)."""

    def __init__(self, slab: PeerSlab, iters: int, B: int, S: int, per: int, sizes: Optional[List[int]] = None):
        self.slab = slab
        self.iters, self.B, self.S, self.per = iters, B, S, per
        if sizes is None:
            self.n_total = per * slab.world
            self.n_offset = per * slab.rank
        else:
            assert len(sizes) == slab.world and sizes[slab.rank] == per
            self.n_total = int(sum(sizes))
            self.n_offset = int(sum(sizes[:slab.rank]))
        self.off_coords = FLAG_WORDS
        self.off_vis = self.off_coords + iters * B * S * self.n_total * 2
        self.off_ffeat = self.off_vis + B * S * self.n_total
        self.words = self.off_ffeat + B * self.n_total * LATENT
        self.key = (slab.generation, slab.local, tuple(slab.ptrs), iters, B, S, per, self.n_total, self.n_offset)

    @staticmethod
    def words_needed(world: int, iters: int, B: int, S: int, per: int, n_total: Optional[int] = None) -> int:
        nt = per * world if n_total is None else n_total
        return FLAG_WORDS + iters * B * S * nt * 2 + B * S * nt + B * nt * LATENT

    def coord_bases(self, it: int) -> List[int]:
        """Address of iteration ``it``'s (B,S,n_total,2) block in every rank's slab."""
        off = 4 * (self.off_coords + it * self.B * self.S * self.n_total * 2)
        return [p + off for p in self.slab.ptrs]

    def views(self):
        """(coords (iters,B,S,n_total,2), vis (B,S,n_total), ffeat (B,n_total,128)) views of the LOCAL slab."""
        w = self.slab.words
        c = w[self.off_coords:self.off_vis].view(self.iters, self.B, self.S, self.n_total, 2)
        v = w[self.off_vis:self.off_ffeat].view(self.B, self.S, self.n_total)
        f = w[self.off_ffeat:self.words].view(self.B, self.n_total, LATENT)
        return c, v, f


def same_host(group) -> bool:
    names: List[Optional[str]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(names, socket.gethostname(), group=group)
    return len(set(names)) == 1
