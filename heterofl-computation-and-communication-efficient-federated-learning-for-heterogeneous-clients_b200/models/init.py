"""Parameter initialisation with PyTorch-default statistics.

The reference relies on ``nn.Conv2d`` / ``nn.Linear`` / ``nn.Embedding`` defaults plus ``init_param``
(``src/models/utils.py:4-10``: norm weight=1 / bias=0, Linear bias=0) and the encoder layer's own
``normal_(0, 0.02)`` on the FFN weights (``src/models/transformer.py:103-110``).  ``nn.TransformerEncoder``
deep-copies ONE layer, so all encoder layers start from identical values -- reproduced here.
"""
from __future__ import annotations

import math
import re
from typing import Optional

import torch

from ..plan.layout import Layout
from ..plan.spec import ModelSpec


def _fan_in(shape) -> int:
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def init_arena(spec: ModelSpec, layout: Layout, arena: torch.Tensor,
               generator: Optional[torch.Generator] = None) -> torch.Tensor:
    views = layout.views(arena)
    by = spec.by_name()
    first_layer = {}
    for name, t in views.items():
        p = by[name]
        kind = p.init
        mlayer = re.match(r"transformer_encoder\.layers\.(\d+)\.(.*)", name)
        if mlayer and int(mlayer.group(1)) > 0:
            t.copy_(first_layer[mlayer.group(2)])
            continue
        if kind in ("conv", "linear"):
            bound = 1.0 / math.sqrt(_fan_in(t.shape))           # kaiming_uniform(a=sqrt(5))
            t.copy_((torch.rand(t.shape, generator=generator) * 2 - 1) * bound)
        elif kind == "conv_bias":
            w = views[name.replace(".bias", ".weight")]
            bound = 1.0 / math.sqrt(_fan_in(w.shape))
            t.copy_((torch.rand(t.shape, generator=generator) * 2 - 1) * bound)
        elif kind == "embedding":
            t.copy_(torch.randn(t.shape, generator=generator))
        elif kind == "normal02":
            t.copy_(torch.randn(t.shape, generator=generator) * 0.02)
        elif kind == "ones":
            t.fill_(1.0)
        elif kind == "zeros":
            t.zero_()
        else:
            raise ValueError(kind)
        if mlayer:
            first_layer[mlayer.group(2)] = t.clone()
    return arena
