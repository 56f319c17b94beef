"""Model factory: a :class:`Model` binds a ModelSpec to one width (rate) and exposes layout + forward."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..ops.nn import NormCtx
from ..plan.layout import Layout
from ..plan.spec import ModelSpec, bn_buffer_names, build_spec
from .init import init_arena
from .transformer import transformer_forward
from .vision import VISION_FORWARD


class Model:
    """A sub-network of ``spec`` at absolute width ``rate`` (``scale_rate = rate / global_rate`` feeds the Scaler,
    reference ``src/models/resnet.py:165``)."""

    def __init__(self, spec: ModelSpec, cfg: Dict, rate: float, channels_last: bool = False):
        self.spec, self.cfg, self.rate = spec, cfg, float(rate)
        self.layout = Layout(spec, rate, channels_last)
        self.scale_rate = (self.rate / float(cfg["global_model_rate"])) if cfg["scale"] else 1.0
        self.is_lm = spec.family == "transformer"

    def new_arena(self, device="cpu", dtype=torch.float32, generator=None, init: bool = True) -> torch.Tensor:
        a = torch.zeros(self.layout.size, dtype=torch.float32)
        if init:
            init_arena(self.spec, self.layout, a, generator)
        return a.to(device=device, dtype=dtype)

    def convert_arena(self, arena: torch.Tensor, other: "Model") -> torch.Tensor:
        """Re-lay ``arena`` (this model's layout) into ``other``'s layout (e.g. host K,C,R,S <-> device K,R,S,C)."""
        out = torch.zeros(other.layout.size, dtype=arena.dtype, device=arena.device)
        src, dst = self.layout.views(arena), other.layout.views(out)
        for k in src:
            dst[k].copy_(src[k])
        return out

    def forward(self, P: Dict[str, torch.Tensor], inp: Dict, ctx: Optional[NormCtx] = None,
                label_mask: Optional[torch.Tensor] = None, want_score: bool = True) -> Dict:
        ctx = ctx if ctx is not None else NormCtx("train")
        if self.is_lm:
            return transformer_forward(self.spec, P, inp, ctx, self.scale_rate, label_mask, want_score,
                                       self.cfg.get("mask_rate", 0.15), self.cfg["transformer"]["dropout"])
        return VISION_FORWARD[self.spec.family](self.spec, P, inp, ctx, self.scale_rate, label_mask, want_score)

    def bn_layers(self):
        """[(name, width)] of BatchNorm layers at this rate."""
        return [(n, ax.local(self.rate)) for n, ax in bn_buffer_names(self.spec)]


def make_model(cfg: Dict, rate: Optional[float] = None, spec: Optional[ModelSpec] = None,
               channels_last: bool = False) -> Model:
    spec = spec if spec is not None else build_spec(cfg)
    return Model(spec, cfg, cfg["global_model_rate"] if rate is None else rate, channels_last)
