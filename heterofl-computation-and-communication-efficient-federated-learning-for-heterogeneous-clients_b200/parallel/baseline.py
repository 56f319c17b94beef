"""The library baseline for aggregation: all-reduce of padded-to-global-width sum and count buffers.

This is what a straightforward ``torch.distributed`` port of reference ``Federation.combine`` does and what the fused
P2P kernel (``parallel/peer.py``) is measured against: every rank scatters its clients into two fp32 buffers of the
GLOBAL model's size (sum, count), two all-reduces move 2 x |W_g| regardless of how narrow the clients were, then
``W_g[count>0] = sum/count``.  It is also the CPU/gloo plumbing path.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

from ..fed.federation import _maps
from ..plan.layout import FedTables


def partial_sums(tables: FedTables, like: torch.Tensor, arenas: Sequence[torch.Tensor], rates: Sequence[float],
                 masks: Sequence[torch.Tensor]):
    """(sum, count) flat buffers shaped like the global arena for the given clients."""
    dev = like.device
    acc = torch.zeros(like.numel(), dtype=torch.float32, device=dev)
    cnt = torch.zeros(like.numel(), dtype=torch.float32, device=dev)
    av, cv = tables.global_layout.views(acc), tables.global_layout.views(cnt)
    for arena, rate, lm in zip(arenas, rates, masks):
        lay = tables.level_layouts[tables.level_id(rate)]
        maps = _maps(tables, rate)
        for p in tables.spec.params:
            loc = lay.view(arena, p.name).to(device=dev, dtype=torch.float32)
            idx = maps[p.name]
            if p.inp is None:
                oi = idx.to(dev)
                if p.label_rows:
                    keep = lm.to(dev)[oi].bool()
                    oi, loc = oi[keep], loc[keep]
                av[p.name][oi] += loc
                cv[p.name][oi] += 1
            else:
                oi, ii = idx[0].to(dev), idx[1].to(dev)
                if p.label_rows:
                    keep = lm.to(dev)[oi].bool()
                    oi, loc = oi[keep], loc[keep]
                av[p.name][oi[:, None], ii[None, :]] += loc
                cv[p.name][oi[:, None], ii[None, :]] += 1
    return acc, cnt


def allreduce_aggregate(tables: FedTables, global_arena: torch.Tensor, levels: List[int], results: Dict[int, Dict],
                        masks_all: List[torch.Tensor], info) -> None:
    order = sorted(results.keys())
    acc, cnt = partial_sums(tables, global_arena, [results[i]["arena"] for i in order],
                            [tables.rates[levels[i]] for i in order], [masks_all[i] for i in order])
    if info.world > 1:
        dist.all_reduce(acc)
        dist.all_reduce(cnt)
    covered = cnt > 0
    global_arena[covered] = (acc[covered] / cnt[covered]).to(global_arena.dtype)
