"""Flat-arena layouts and index maps derived from a :class:`ModelSpec`.

Every model instance (the global model and each client sub-network) stores all of its parameters in ONE
contiguous buffer (``arena``); a :class:`Layout` records where each tensor lives.  Because every slice in
HeteroFL is a prefix (or a per-head prefix) the map local-coordinate -> global-coordinate of a tensor is fully
described by 8 small integers per (tensor, level) -- the ``SliceDesc`` rows consumed by the fused
distribute / aggregate kernels (``csrc/fed_kernels.cu``) -- instead of the materialised ``arange`` index tensors
of the reference (``src/fed.py:26-159``).  :func:`index_maps` still produces those tensors for the CPU fallback
and for oracle tests against ``Federation.split_model``.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .spec import FIXED, HEADS, PREFIX, Axis, ModelSpec

ALIGN = 8  # elements; keeps every tensor 32 B (fp32) / 16 B (bf16) aligned inside the arena


@dataclass(frozen=True)
class Entry:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int


class Layout:
    """Offsets of every parameter of a model built at absolute width ``rate``."""

    def __init__(self, spec: ModelSpec, rate: float, channels_last: bool = False):
        self.spec, self.rate, self.channels_last = spec, float(rate), bool(channels_last)
        self.entries: "OrderedDict[str, Entry]" = OrderedDict()
        off = 0
        for p in spec.params:
            shp = p.shape(rate)
            n = int(np.prod(shp))
            self.entries[p.name] = Entry(p.name, shp, off, n)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.size = off                      # arena length in elements (padded)
        self.num_params = sum(e.numel for e in self.entries.values())

    def view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        """Zero-copy view of one parameter.  With ``channels_last`` (device arenas) 4-D conv filters are *stored*
        K,R,S,C -- the GEMM-K axis (r,s,c) is contiguous, which is what NHWC implicit-GEMM kernels and cuDNN's NHWC
        path consume without a per-step re-layout -- and returned as a logical ``[K, C, R, S]`` tensor with
        channels-last strides.  Host arenas keep the plain K,C,R,S order."""
        e = self.entries[name]
        flat = arena[e.offset:e.offset + e.numel]
        if len(e.shape) == 4 and self.channels_last:
            K, C, R, S = e.shape
            return flat.view(K, R, S, C).permute(0, 3, 1, 2)
        return flat.view(e.shape)

    def views(self, arena: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        """Zero-copy per-parameter views into a flat arena tensor."""
        return OrderedDict((name, self.view(arena, name)) for name in self.entries)

    def state_dict(self, arena: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        """Reference-format state dict (contiguous ``[K, C, R, S]`` filters)."""
        return OrderedDict((k, v.detach().clone(memory_format=torch.contiguous_format))
                           for k, v in self.views(arena).items())

    def load_state_dict(self, arena: torch.Tensor, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        v = self.views(arena)
        for k, t in v.items():
            if k in sd:
                t.copy_(sd[k].reshape(t.shape))
            elif strict:
                raise KeyError("missing key %s" % k)


# ----------------------------------------------------------------------------------------------------------
# index maps (reference-compatible tensors; CPU fallback + oracle tests)
# ----------------------------------------------------------------------------------------------------------
def axis_index(ax: Axis, rate: float, global_rate: float) -> torch.Tensor:
    """Global indices (into the global model's axis) kept by a client at ``rate``."""
    g = ax.local(global_rate)
    if ax.kind == FIXED:
        return torch.arange(g)
    if ax.kind == PREFIX:
        return torch.arange(g)[:ax.local(rate)]
    dg, dl = g // ax.heads, ax.local(rate) // ax.heads
    return torch.arange(g).reshape(ax.heads, dg)[:, :dl].reshape(-1)


def index_maps(spec: ModelSpec, rate: float, global_rate: float) -> "OrderedDict[str, object]":
    """``{name: (out_idx, in_idx)}`` for >=2-D tensors, ``{name: idx}`` for vectors -- what
    ``Federation.split_model`` (reference ``src/fed.py:26-159``) returns for one client."""
    out = OrderedDict()
    for p in spec.params:
        if p.inp is None:
            out[p.name] = axis_index(p.out, rate, global_rate)
        else:
            out[p.name] = (axis_index(p.out, rate, global_rate), axis_index(p.inp, rate, global_rate))
    return out


# ----------------------------------------------------------------------------------------------------------
# kernel descriptor tables
# ----------------------------------------------------------------------------------------------------------
# One row per (tensor):   [g_off, R, C, inner, rH, cH, label_rows, col_innermost]
#   physical order of a tensor is [R, C, inner] (col_innermost=0) or [R, inner, C] (col_innermost=1, conv filters)
# One row per (tensor, level): [l_off, r_loc, c_loc, r_dl, r_dg, c_dl, c_dg, _pad]
#   local row i = h*r_dl + t  <->  global row h*r_dg + t   (rH heads; plain prefix: rH=1, r_dl=r_loc, r_dg=R)
TENSOR_COLS = 8
LEVEL_COLS = 8


def _axis_desc(ax: Optional[Axis], rate: float, global_rate: float) -> Tuple[int, int, int, int, int]:
    """(global size, local size, heads, d_local, d_global)"""
    if ax is None:
        return 1, 1, 1, 1, 1
    g, l = ax.local(global_rate), ax.local(rate)
    if ax.kind == HEADS:
        return g, l, ax.heads, l // ax.heads, g // ax.heads
    return g, l, 1, l, g


@dataclass
class FedTables:
    """Device-consumable description of how L levels nest inside the global arena."""
    spec: ModelSpec
    global_rate: float
    rates: List[float]                   # distinct client rates, index = level id
    global_layout: Layout
    level_layouts: List[Layout]
    tensor_tab: np.ndarray               # [T, TENSOR_COLS] int64
    level_tab: np.ndarray                # [L, T, LEVEL_COLS] int64

    def level_id(self, rate: float) -> int:
        for i, r in enumerate(self.rates):
            if abs(r - rate) < 1e-12:
                return i
        raise KeyError(rate)


def build_tables(spec: ModelSpec, global_rate: float, rates: Sequence[float],
                 channels_last: bool = False) -> FedTables:
    rates = sorted({float(r) for r in rates}, reverse=True)
    gl = Layout(spec, global_rate, channels_last)
    lls = [Layout(spec, r, channels_last) for r in rates]
    T = len(spec.params)
    tt = np.zeros((T, TENSOR_COLS), dtype=np.int64)
    lt = np.zeros((len(rates), T, LEVEL_COLS), dtype=np.int64)
    for t, p in enumerate(spec.params):
        R, _, rH, _, _ = _axis_desc(p.out, global_rate, global_rate)
        C, _, cH, _, _ = _axis_desc(p.inp, global_rate, global_rate)
        tt[t] = [gl.entries[p.name].offset, R, C, p.inner, rH, cH, int(p.label_rows),
                 int(len(p.tail) == 2 and channels_last)]
        for li, r in enumerate(rates):
            _, rl, _, rdl, rdg = _axis_desc(p.out, r, global_rate)
            _, cl, _, cdl, cdg = _axis_desc(p.inp, r, global_rate)
            lt[li, t] = [lls[li].entries[p.name].offset, rl, cl, rdl, rdg, cdl, cdg, 0]
    return FedTables(spec, float(global_rate), list(rates), gl, lls, tt, lt)
