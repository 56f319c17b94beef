"""Process-group plumbing: one process per GPU, ``torch.distributed`` with NCCL (GPU) or gloo (CPU).

The reference has no distributed mode at all; the only multi-device code is ``nn.DataParallel`` in the
centralised trainers (``src/train_classifier.py:65-66``).  Here the federated round itself is sharded over ranks.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


@dataclass
class DistInfo:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    device: torch.device = torch.device("cpu")
    backend: Optional[str] = None

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def init_distributed(device: Optional[str] = None, timeout_s: int = 600) -> DistInfo:
    """Read RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (torchrun contract) and create the default process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    use_cuda = torch.cuda.is_available() and (device is None or str(device).startswith("cuda"))
    dev = torch.device("cuda", local % max(torch.cuda.device_count(), 1)) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(dev)
    backend = None
    if world > 1:
        backend = "nccl" if use_cuda else "gloo"
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            kw = dict(device_id=dev) if use_cuda else {}
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
    return DistInfo(rank, world, local, dev, backend)


def barrier(info: DistInfo) -> None:
    if info.world > 1:
        dist.barrier()


_PEER_POOL = None


def set_peer_pool(pool) -> None:
    """Route small all-reduces through the peer-memory kernel path (``parallel/peer.py``) instead of NCCL."""
    global _PEER_POOL
    _PEER_POOL = pool


import threading

_TLS = threading.local()


def _channel() -> int:
    return getattr(_TLS, "channel", 0)


class peer_channel:
    """Context manager (thread-local): small all-reduces issued inside use barrier channel ``ch`` -- one channel per
    concurrently used CUDA stream / host thread."""

    def __init__(self, ch: int):
        self.ch, self.prev = ch, 0

    def __enter__(self):
        self.prev = _channel()
        _TLS.channel = self.ch
        return self

    def __exit__(self, *a):
        _TLS.channel = self.prev


def all_reduce_sum(info: DistInfo, t: torch.Tensor) -> torch.Tensor:
    if info.world > 1:
        if _PEER_POOL is not None and t.is_cuda and t.numel() <= _PEER_POOL.SCRATCH_FLOATS \
                and t.dtype in (torch.float32, torch.float64):
            _PEER_POOL.all_reduce_sum_(t, _channel())       # fp64 meters travel as fp32 (sums < 1e6, counts exact)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_max(info: DistInfo, t: torch.Tensor) -> torch.Tensor:
    if info.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def broadcast(info: DistInfo, t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if info.world > 1:
        dist.broadcast(t, src)
    return t


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
