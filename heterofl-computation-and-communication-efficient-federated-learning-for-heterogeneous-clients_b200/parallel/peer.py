"""Peer-mapped memory over NVLink / NVSwitch and the fused multi-GPU aggregation.

Every rank owns one ``cudaMalloc`` pool that all other ranks map through CUDA IPC (``csrc/bindings.cpp``:
``peer_alloc / ipc_get_handle / ipc_open_handle``).  Pool layout (identical on every rank, so offsets are computed,
never communicated)::

    [ flags: uint32[world] | persistent: the W_g replica | round: finished client arenas (bump-allocated) ]

Aggregation of a round (``p2p_aggregate``) is then

    barrier_kernel  ->  aggregate_kernel  ->  barrier_kernel          (all on the compute stream, no host sync)

where the aggregate kernel of rank r handles block-shard r of the global coordinates, *loads* the covering clients'
values straight from the owning GPUs' pools (P2P reads over NVSwitch), divides by the on-the-fly coverage count and
*stores* the result into the W_g replica of every rank (P2P writes) -- gather + reduce + all-gather in one pass.
Bytes on the wire per round = sum of client level sizes (pull) + |W_g| (publish) instead of 2 x 2 x |W_g| for
the padded sum/count all-reduce baseline (``parallel/baseline.py``).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops import cuda as cu
from ..plan.layout import FedTables

_ALIGN = 256
_CHANNELS = 4                      # independent barrier / all-reduce channels (one per concurrently used stream)
_FLAG_STRIDE = 256                 # bytes per channel flag row (>= 4 * world)
_FLAG_BYTES = _CHANNELS * _FLAG_STRIDE


def _align(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class PeerPool:
    def __init__(self, info, nbytes: int):
        assert info.device.type == "cuda"
        self.info = info
        self.nbytes = _align(int(nbytes) + _FLAG_BYTES)
        self.C = cu.C()
        self.base = int(self.C.peer_alloc(self.nbytes))
        self.off = _FLAG_BYTES
        self.round_base = self.off
        self.peer_base: List[int] = [0] * info.world
        self.peer_base[info.rank] = self.base
        self.epoch = [0] * _CHANNELS
        self.err = torch.zeros(1, dtype=torch.int32, device=info.device)
        self._flag_ptrs: List[Optional[torch.Tensor]] = [None] * _CHANNELS
        self._opened: List[int] = []

    # -- allocation (deterministic bump allocator) -----------------------------------------------------------
    def _alloc(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        nb = _align(numel * torch.empty(0, dtype=dtype).element_size())
        if self.off + nb > self.nbytes:
            raise MemoryError("peer pool exhausted (%d + %d > %d)" % (self.off, nb, self.nbytes))
        name = {torch.float32: "float32", torch.bfloat16: "bfloat16", torch.uint8: "uint8", torch.int32: "int32",
                torch.int64: "int64", torch.float64: "float64"}[dtype]
        t = self.C.tensor_from_ptr(self.base + self.off, int(numel), name, self.info.device.index or 0)
        t._hfl_pool_offset = self.off
        self.off += nb
        return t

    def alloc_persistent(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        t = self._alloc(numel, dtype)
        self.round_base = self.off
        return t

    def alloc_round(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        return self._alloc(numel, dtype)

    def reset_round(self) -> None:
        self.off = self.round_base

    @staticmethod
    def round_offsets(round_base: int, sizes_bytes: List[int]) -> List[int]:
        out, off = [], round_base
        for nb in sizes_bytes:
            out.append(off)
            off += _align(nb)
        return out

    # -- handle exchange ---------------------------------------------------------------------------------------
    def exchange(self) -> None:
        handle = bytes(self.C.ipc_get_handle(self.base))
        handles: List = [None] * self.info.world
        dist.all_gather_object(handles, handle)
        for r, h in enumerate(handles):
            if r == self.info.rank:
                continue
            p = int(self.C.ipc_open_handle(h))
            self.peer_base[r] = p
            self._opened.append(p)
        for ch in range(_CHANNELS):
            self._flag_ptrs[ch] = torch.tensor([b + ch * _FLAG_STRIDE for b in self.peer_base],
                                               dtype=torch.int64).to(self.info.device)
        dist.barrier()

    def barrier(self, channel: int = 0) -> None:
        """Stream-ordered cross-GPU barrier (device-side flags, no host sync).  Streams that run concurrently must use
        different channels (every rank has to issue the barriers of one channel in the same order)."""
        self.epoch[channel] += 1
        self.C.peer_barrier(self._flag_ptrs[channel], self.info.rank, self.info.world, self.epoch[channel], self.err)

    # -- small one-shot all-reduce over peer memory (sBN statistic merge, metric merge) ----------------------------
    SCRATCH_FLOATS = 1 << 16

    def enable_all_reduce(self) -> None:
        """Reserve the persistent scratch slabs (must be called on every rank BEFORE ``exchange``)."""
        self._scratch = [self.alloc_persistent(self.SCRATCH_FLOATS, torch.float32) for _ in range(_CHANNELS)]
        self._scratch_ptrs = [None] * _CHANNELS

    def all_reduce_sum_(self, t: torch.Tensor, channel: int = 0) -> torch.Tensor:
        """In-place sum of a small fp32/fp64 tensor over all ranks: every rank publishes its values in its pool slab,
        a flag barrier orders the stores, ``peer_sum_kernel`` reads all slabs over NVLink, a second barrier frees the
        slab for the next call.  No NCCL, no host synchronisation."""
        n = t.numel()
        assert n <= self.SCRATCH_FLOATS, "all_reduce_sum_ is meant for small vectors"
        if self._scratch_ptrs[channel] is None:
            off = self._scratch[channel]._hfl_pool_offset
            self._scratch_ptrs[channel] = torch.tensor([b + off for b in self.peer_base],
                                                       dtype=torch.int64).to(self.info.device)
        src = t.reshape(-1)
        self._scratch[channel][:n].copy_(src.to(torch.float32))
        self.barrier(channel)
        out = torch.empty(n, dtype=torch.float32, device=self.info.device)
        self.C.peer_sum(self._scratch_ptrs[channel], out)
        self.barrier(channel)
        src.copy_(out.to(t.dtype))
        return t

    def check(self) -> None:
        if int(self.err.item()) != 0:
            raise RuntimeError("peer barrier timed out (a rank did not arrive)")

    def close(self) -> None:
        for p in self._opened:
            try:
                self.C.ipc_close_handle(p)
            except Exception:
                pass
        self._opened = []


def p2p_aggregate(pool: PeerPool, tables: FedTables, global_arena: torch.Tensor, levels: List[int],
                  assignment: List[List[int]], masks_all: List[torch.Tensor], info, active=None,
                  reduce: Optional[torch.Tensor] = None) -> None:
    """``assignment[r]`` = client positions trained by rank r, in the order their result arenas were allocated;
    ``active`` (optional) restricts the average to a subset of the clients (fault injection / stragglers);
    ``reduce`` (optional, small fp32/fp64 vector) is summed over all ranks in place by the same launch."""
    from ..ops import fed_ops
    dev = global_arena.device
    dt = fed_ops.device_tables(tables, dev)
    M = len(levels)
    ptrs = [0] * M
    for r, mine in enumerate(assignment):
        sizes = [tables.level_layouts[levels[i]].size * 4 for i in mine]
        for i, off in zip(mine, PeerPool.round_offsets(pool.round_base, sizes)):
            ptrs[i] = pool.peer_base[r] + off
    if active is not None:
        ptrs, levels, masks_all = ([ptrs[i] for i in active], [levels[i] for i in active],
                                   [masks_all[i] for i in active])
    g_off = global_arena._hfl_pool_offset
    out_ptr = torch.tensor([b + g_off for b in pool.peer_base], dtype=torch.int64).to(dev, non_blocking=True)
    nb = dt.n_global_blocks2
    lo, hi = info.rank * nb // info.world, (info.rank + 1) * nb // info.world
    red_src = red_dst = None
    if reduce is not None:
        # the round's metric vector rides along: published into this rank's slab before the barrier, summed over all
        # peers' slabs by block 0 of the aggregation launch (no peer_sum launch, no extra barrier pair)
        n = reduce.numel()
        ch = 3                                        # dedicated slab: never used by the evaluation streams
        if pool._scratch_ptrs[ch] is None:
            off = pool._scratch[ch]._hfl_pool_offset
            pool._scratch_ptrs[ch] = torch.tensor([b + off for b in pool.peer_base], dtype=torch.int64).to(dev)
        pool._scratch[ch][:n].copy_(reduce.reshape(-1).to(torch.float32))
        red_src = pool._scratch_ptrs[ch]
        red_dst = torch.empty(n, dtype=torch.float32, device=dev)
    pool.barrier()                       # every rank's finished client arenas (and metric slab) are in place
    fed_ops.aggregate_ptrs(dt, ptrs, levels, masks_all, out_ptr, lo, hi - lo, red_src, red_dst)
    pool.barrier()                       # every shard of the new W_g has been published everywhere
    if reduce is not None:
        reduce.copy_(red_dst.to(reduce.dtype).view_as(reduce))
