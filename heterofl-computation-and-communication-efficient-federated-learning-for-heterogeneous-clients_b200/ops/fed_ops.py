"""Device-side distribute / aggregate: descriptor tables + launch wrappers for ``csrc/fed_kernels.cu``."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from ..plan.layout import FedTables
from . import cuda as cu


class DeviceTables:
    """FedTables uploaded once per (tables, device): tensor/level descriptors + block work lists."""

    def __init__(self, tables: FedTables, device):
        self.tables, self.device = tables, device
        chunk = cu.C().fed_chunk()
        self.chunk = chunk
        T = tables.tensor_tab.shape[0]
        self.tensor_tab = torch.from_numpy(tables.tensor_tab.astype(np.int32)).to(device)
        self.level_tab = torch.from_numpy(tables.level_tab.astype(np.int32).reshape(-1, 8)).to(device)
        # aggregate: blocks over the GLOBAL tensors
        gb = []
        for t in range(T):
            _, R, C, inner = tables.tensor_tab[t, :4]
            n = int(R * C * inner)
            gb += [(t, s) for s in range(0, n, chunk)]
        self.n_global_blocks = len(gb)
        self.global_blocks = torch.tensor(gb, dtype=torch.int32).reshape(-1, 2).to(device)
        # aggregate v2: bigger chunks + per-tensor "4 consecutive innermost coordinates behave alike" flag
        chunk2 = cu.C().fed2_chunk()
        gb2, vec = [], []
        L = len(tables.rates)
        for t in range(T):
            _, R, C, inner, rH, cH, label_rows, col_inner = [int(v) for v in tables.tensor_tab[t]]
            n = R * C * inner
            gb2 += [(t, s) for s in range(0, n, chunk2)]
            lv = tables.level_tab[:, t, :]                         # [L, 8]: l_off r_loc c_loc r_dl r_dg c_dl c_dg
            if C > 1 or inner > 1:                                 # innermost axis = columns (K,R,S,C filters / matrices)
                inner_ok = (col_inner == 1 or inner == 1)
                ok = inner_ok and C % 4 == 0 and all(int(lv[i, 2]) % 4 == 0 for i in range(L)) and \
                    (cH == 1 or all(int(lv[i, 5]) % 4 == 0 and int(lv[i, 6]) % 4 == 0 for i in range(L)))
            else:                                                  # vectors: innermost axis = rows
                ok = (not label_rows) and R % 4 == 0 and all(int(lv[i, 1]) % 4 == 0 for i in range(L)) and \
                    (rH == 1 or all(int(lv[i, 3]) % 4 == 0 and int(lv[i, 4]) % 4 == 0 for i in range(L)))
            vec.append(1 if ok else 0)
        self.n_global_blocks2 = len(gb2)
        self.global_blocks2 = torch.tensor(gb2, dtype=torch.int32).reshape(-1, 2).to(device)
        self.vec4 = torch.tensor(vec, dtype=torch.int32).to(device)
        self.n_levels = L
        # distribute: blocks over each level's LOCAL tensors
        lb, starts = [], [0]
        for li in range(len(tables.rates)):
            for t in range(T):
                inner = int(tables.tensor_tab[t, 3])
                n = int(tables.level_tab[li, t, 1] * tables.level_tab[li, t, 2] * inner)
                lb += [(t, s) for s in range(0, n, chunk)]
            starts.append(len(lb))
        self.level_blocks = torch.tensor(lb, dtype=torch.int32).reshape(-1, 2).to(device)
        self.level_block_start = torch.tensor(starts, dtype=torch.int32).to(device)
        self.level_block_counts = [starts[i + 1] - starts[i] for i in range(len(tables.rates))]


def device_tables(tables: FedTables, device) -> DeviceTables:
    cache = tables.__dict__.setdefault("_dev_tables", {})
    key = str(device)
    if key not in cache:
        cache[key] = DeviceTables(tables, device)
    return cache[key]


def _ptr_tensor(ts: Sequence[torch.Tensor], device) -> torch.Tensor:
    return torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64).to(device, non_blocking=True)


def distribute(tables: FedTables, global_arena: torch.Tensor, w: Sequence[torch.Tensor], levels: Sequence[int],
               wc: Optional[Sequence[torch.Tensor]] = None, mom: Optional[Sequence[torch.Tensor]] = None) -> None:
    """One launch: slice ``W_g`` into every client's fp32 master arena (+ bf16 copy, + zeroed momentum)."""
    if len(w) == 0:
        return
    dev = global_arena.device
    dt = device_tables(tables, dev)
    lv = torch.tensor(list(levels), dtype=torch.int32).to(dev, non_blocking=True)
    max_blocks = max(dt.level_block_counts[l] for l in levels)
    cu.C().distribute(dt.tensor_tab, dt.level_tab, dt.level_blocks, dt.level_block_start, max_blocks, global_arena,
                      _ptr_tensor(w, dev), _ptr_tensor(wc, dev) if wc is not None else None,
                      _ptr_tensor(mom, dev) if mom is not None else None, lv)


def aggregate(tables: FedTables, global_arena: torch.Tensor, client_arenas: Sequence[torch.Tensor],
              levels: Sequence[int], label_masks: Sequence[torch.Tensor]) -> None:
    """One launch: heterogeneous FedAvg of local client arenas into the (single, local) global arena."""
    dev = global_arena.device
    dt = device_tables(tables, dev)
    aggregate_ptrs(dt, [int(a.data_ptr()) for a in client_arenas], list(levels), label_masks,
                   _ptr_tensor([global_arena], dev), 0, dt.n_global_blocks2)


def aggregate_ptrs(dt: DeviceTables, client_ptrs: Sequence[int], levels: Sequence[int],
                   label_masks: Sequence[torch.Tensor], out_ptr: torch.Tensor, block_lo: int, n_blocks: int,
                   red_src: Optional[torch.Tensor] = None, red_dst: Optional[torch.Tensor] = None) -> None:
    """Vectorised aggregation (``aggregate_v2_kernel``): clients are sorted by level on the host (stable, so the
    fp32 summation order inside a level is the client order); any number of clients; pointers may be peer-mapped.
    ``red_src`` / ``red_dst``: optional small all-reduce (peer slab pointers -> local fp32 vector) in the same launch."""
    dev = dt.device
    order = sorted(range(len(levels)), key=lambda i: levels[i])
    L = dt.n_levels
    starts = [0] * (L + 1)
    for lv in levels:
        starts[lv + 1] += 1
    for i in range(L):
        starts[i + 1] += starts[i]
    cp = torch.tensor([client_ptrs[i] for i in order], dtype=torch.int64).to(dev, non_blocking=True)
    ls = torch.tensor(starts, dtype=torch.int32).to(dev, non_blocking=True)
    if len(order):
        masks = torch.stack([label_masks[i].to(torch.uint8) for i in order]).contiguous().to(dev, non_blocking=True)
    else:
        masks = torch.zeros(0, 1, dtype=torch.uint8, device=dev)
    cu.C().aggregate_v2(dt.tensor_tab, dt.level_tab, dt.vec4, dt.global_blocks2, int(block_lo), int(n_blocks), cp, ls,
                        masks, out_ptr, red_src, red_dst)


def aggregate_raw(dt: DeviceTables, client_ptr: torch.Tensor, client_level: torch.Tensor, masks: torch.Tensor,
                  out_ptr: torch.Tensor, block_lo: int, n_blocks: int) -> None:
    """Pointer-level entry used by the multi-GPU path (clients / outputs may be peer-mapped)."""
    M = client_ptr.numel()
    maxc = cu.C().fed_max_clients()
    if M > maxc:
        raise ValueError("aggregate supports at most %d clients per launch, got %d" % (maxc, M))
    cu.C().aggregate(dt.tensor_tab, dt.level_tab, dt.global_blocks, int(block_lo), int(n_blocks), client_ptr,
                     client_level, masks, out_ptr)
