"""Width-parameterised vision models as *functions of a parameter dict*.

A model is not an ``nn.Module`` owning tensors: its parameters are views into a flat arena
(:mod:`hfl_b200.plan.layout`) handed to ``forward`` -- the same code runs the global model and every client
sub-network; only shapes differ.  Architecture parity: MNIST CNN = reference ``src/models/conv.py:10-82``;
pre-activation ResNet-18/34/50/101/152 = ``src/models/resnet.py:10-208``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from ..ops.nn import NormCtx, masked_cross_entropy, masked_logits, norm_act
from ..plan.spec import ModelSpec


def _out(score: torch.Tensor, inp: Dict, label_mask: Optional[torch.Tensor], want_score: bool) -> Dict:
    if want_score == "raw":          # evaluation: unmasked logits only (masks / metrics applied by the caller)
        return {"logits": score.float()}
    out = {"loss": masked_cross_entropy(score, inp["label"], label_mask, inp.get("_meter"))}
    if want_score:
        out["score"] = masked_logits(score.float(), label_mask)
    return out


def conv_forward(spec: ModelSpec, P: Dict[str, torch.Tensor], inp: Dict, ctx: NormCtx, scale_rate: float,
                 label_mask: Optional[torch.Tensor] = None, want_score: bool = True) -> Dict:
    """4 x [conv3x3+bias -> Scaler -> norm -> ReLU -> maxpool2] (last pool dropped) -> avgpool -> linear."""
    norm, hidden = spec.meta["norm"], spec.meta["hidden"]
    x = inp["img"]
    nb = len(hidden)
    for i in range(nb):
        b = 5 * i
        x = F.conv2d(x, P["blocks.%d.weight" % b], P["blocks.%d.bias" % b], 1, 1)
        pre = "blocks.%d" % (b + 2)
        x = norm_act(x, norm, P.get(pre + ".weight"), P.get(pre + ".bias"), pre, ctx, scale_rate)
        if i < nb - 1:
            x = F.max_pool2d(x, 2)
    x = x.mean(dim=(2, 3))
    lin = 5 * nb + 1
    score = F.linear(x, P["blocks.%d.weight" % lin], P["blocks.%d.bias" % lin])
    return _out(score, inp, label_mask, want_score)


def resnet_forward(spec: ModelSpec, P: Dict[str, torch.Tensor], inp: Dict, ctx: NormCtx, scale_rate: float,
                   label_mask: Optional[torch.Tensor] = None, want_score: bool = True) -> Dict:
    norm = spec.meta["norm"]

    def na(t, pre):
        return norm_act(t, norm, P.get(pre + ".weight"), P.get(pre + ".bias"), pre, ctx, scale_rate)

    out = F.conv2d(inp["img"], P["conv1.weight"], None, 1, 1)
    for blk in spec.meta["blocks"]:
        pre, stride = blk["prefix"], blk["stride"]
        o = na(out, pre + ".n1")
        sc = F.conv2d(o, P[pre + ".shortcut.weight"], None, stride) if blk["shortcut"] else out
        if blk["kind"] == "block":
            o = F.conv2d(o, P[pre + ".conv1.weight"], None, stride, 1)
            o = F.conv2d(na(o, pre + ".n2"), P[pre + ".conv2.weight"], None, 1, 1)
        else:
            o = F.conv2d(o, P[pre + ".conv1.weight"], None, 1, 0)
            o = F.conv2d(na(o, pre + ".n2"), P[pre + ".conv2.weight"], None, stride, 1)
            o = F.conv2d(na(o, pre + ".n3"), P[pre + ".conv3.weight"], None, 1, 0)
        out = o + sc
    out = na(out, "n4")
    out = out.mean(dim=(2, 3))
    score = F.linear(out, P["linear.weight"], P["linear.bias"])
    return _out(score, inp, label_mask, want_score)


VISION_FORWARD = {"conv": conv_forward, "resnet": resnet_forward}
