"""Server side of HeteroFL: rate assignment, sliced distribution, heterogeneous FedAvg.

Semantics follow reference ``src/fed.py`` (``make_model_rate`` :15-24, ``distribute`` :161-178, ``combine``
:180-297) but the data model is different: the global model and every client sub-network are flat arenas
(:mod:`hfl_b200.plan.layout`), clients of one level share one layout, and both hot paths run as ONE fused kernel
launch over all tensors and all clients (``ops.fed_ops``) -- the pure-torch code in this file is the CPU/gloo
fallback and the oracle the kernels are tested against.

Aggregation rule: every coordinate of ``W_g`` becomes the mean over the clients whose slice covers it (label rows
of the classifier / token-embedding / vocabulary projection additionally need the row to be in the client's label
set); coordinates nobody covers keep their value.  Summation is fp32 in client order.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ..plan.layout import FedTables, Layout, build_tables, index_maps
from ..plan.spec import ModelSpec


def label_mask_tensor(label_split_u: Sequence[int], n_rows: int) -> torch.Tensor:
    """uint8 row mask of a client's label set (rows >= the label range -- e.g. the [MASK] embedding row --
    are never set, so they are never aggregated; reference quirk SURVEY 2.2(i))."""
    m = torch.zeros(n_rows, dtype=torch.uint8)
    idx = torch.as_tensor(list(label_split_u), dtype=torch.long)
    if idx.numel():
        m[idx] = 1
    return m


class Federation:
    def __init__(self, cfg: Dict, spec: ModelSpec, global_arena: torch.Tensor, label_split,
                 backend: str = "auto", channels_last: bool = False):
        self.cfg, self.spec = cfg, spec
        self.global_arena = global_arena
        self.label_split = label_split
        self.global_rate = float(cfg["global_model_rate"])
        levels = sorted(set(float(r) for r in cfg["model_rate"]) | {self.global_rate}, reverse=True)
        self.tables: FedTables = build_tables(spec, self.global_rate, levels, channels_last)
        assert self.tables.global_layout.size == global_arena.numel(), "global arena does not match layout"
        self.backend = backend
        self.label_rows = max([int(p.out.local(self.global_rate)) for p in spec.params if p.label_rows] + [1])
        self.model_rate: np.ndarray = np.zeros(0)
        self.make_model_rate()

    # -- rate assignment ---------------------------------------------------------------------------------
    def make_model_rate(self) -> None:
        cfg = self.cfg
        if cfg["model_split_mode"] == "dynamic":
            # same RNG call as the reference so seeds reproduce its draws (src/fed.py:16-19)
            rate_idx = torch.multinomial(torch.tensor(cfg["proportion"]), num_samples=cfg["num_users"],
                                         replacement=True).tolist()
            self.model_rate = np.array(cfg["model_rate"])[rate_idx]
        elif cfg["model_split_mode"] == "fix":
            self.model_rate = np.array(cfg["model_rate"])
        else:
            raise ValueError("Not valid model split mode")

    def sample_users(self) -> List[int]:
        """``randperm(num_users)[:ceil(frac*N)]`` (reference ``src/train_classifier_fed.py:173-174``)."""
        n = self.cfg["num_users"]
        k = int(np.ceil(self.cfg["frac"] * n))
        return torch.arange(n)[torch.randperm(n)[:k]].tolist()

    def user_label_mask(self, user: int) -> torch.Tensor:
        return label_mask_tensor(self.label_split[user], self.label_rows)

    # -- hot path 1: distribute ---------------------------------------------------------------------------
    def new_client_arena(self, rate: float, device=None, dtype=torch.float32) -> torch.Tensor:
        lay = self.tables.level_layouts[self.tables.level_id(rate)]
        return torch.zeros(lay.size, dtype=dtype, device=device if device is not None else self.global_arena.device)

    def distribute(self, user_idx: Sequence[int], out: Optional[List[torch.Tensor]] = None,
                   redraw: bool = True) -> List[torch.Tensor]:
        """Slice ``W_g`` down to each active client's width.  Returns one flat arena per client (``out`` lets the
        caller supply pre-allocated / peer-mapped arenas)."""
        if redraw:
            self.make_model_rate()
        rates = [float(self.model_rate[u]) for u in user_idx]
        if out is None:
            out = [self.new_client_arena(r) for r in rates]
        if self._use_cuda():
            from ..ops import fed_ops
            fed_ops.distribute(self.tables, self.global_arena, out, [self.tables.level_id(r) for r in rates])
        else:
            distribute_torch(self.tables, self.global_arena, out, rates)
        return out

    # -- hot path 2: combine ------------------------------------------------------------------------------
    def combine(self, client_arenas: Sequence[torch.Tensor], user_idx: Sequence[int]) -> None:
        rates = [float(self.model_rate[u]) for u in user_idx]
        masks = [self.user_label_mask(u) for u in user_idx]
        if self._use_cuda():
            from ..ops import fed_ops
            fed_ops.aggregate(self.tables, self.global_arena, list(client_arenas),
                              [self.tables.level_id(r) for r in rates], masks)
        else:
            combine_torch(self.tables, self.global_arena, client_arenas, rates, masks)

    def _use_cuda(self) -> bool:
        if self.backend == "torch":
            return False
        return self.global_arena.is_cuda


# ----------------------------------------------------------------------------------------------------------
# pure-torch fallback / oracle
# ----------------------------------------------------------------------------------------------------------
def _maps(tables: FedTables, rate: float):
    key = ("_maps", rate)
    cache = tables.__dict__.setdefault("_map_cache", {})
    if key not in cache:
        cache[key] = index_maps(tables.spec, rate, tables.global_rate)
    return cache[key]


def distribute_torch(tables: FedTables, global_arena: torch.Tensor, out: Sequence[torch.Tensor],
                     rates: Sequence[float]) -> None:
    gviews = tables.global_layout.views(global_arena)
    for arena, rate in zip(out, rates):
        lay: Layout = tables.level_layouts[tables.level_id(rate)]
        lviews = lay.views(arena)
        maps = _maps(tables, rate)
        for p in tables.spec.params:
            g, idx = gviews[p.name], maps[p.name]
            if p.inp is None:
                lviews[p.name].copy_(g[idx.to(g.device)])
            else:
                oi, ii = idx[0].to(g.device), idx[1].to(g.device)
                lviews[p.name].copy_(g[oi][:, ii])


def combine_torch(tables: FedTables, global_arena: torch.Tensor, client_arenas: Sequence[torch.Tensor],
                  rates: Sequence[float], label_masks: Sequence[torch.Tensor]) -> None:
    gviews = tables.global_layout.views(global_arena)
    dev = global_arena.device
    for p in tables.spec.params:
        g = gviews[p.name]
        acc = torch.zeros(g.shape, dtype=torch.float32, device=dev)
        cnt = torch.zeros(g.shape, dtype=torch.float32, device=dev)
        for arena, rate, lm in zip(client_arenas, rates, label_masks):
            lay = tables.level_layouts[tables.level_id(rate)]
            loc = lay.view(arena, p.name).to(device=dev, dtype=torch.float32)
            idx = _maps(tables, rate)[p.name]
            if p.inp is None:
                oi = idx.to(dev)
                if p.label_rows:
                    keep = lm.to(dev)[oi].bool()
                    oi, loc = oi[keep], loc[keep]
                acc[oi] += loc
                cnt[oi] += 1
            else:
                oi, ii = idx[0].to(dev), idx[1].to(dev)
                if p.label_rows:
                    keep = lm.to(dev)[oi].bool()
                    oi, loc = oi[keep], loc[keep]
                acc[oi[:, None], ii[None, :]] += loc
                cnt[oi[:, None], ii[None, :]] += 1
        covered = cnt > 0
        g[covered] = (acc[covered] / cnt[covered]).to(g.dtype)
