"""MNIST CNN train step with a hand-scheduled backward (the conv-family twin of :mod:`.fused_resnet`).

Network (reference ``src/models/conv.py:29-82``): ``4 x [conv3x3 + bias -> Scaler -> BatchNorm -> ReLU -> MaxPool2d(2)]``
(no pool after the last block) ``-> AdaptiveAvgPool -> Linear -> masked cross-entropy``.  Fused step:

* ``bn_act_fwd / bn_act_bwd`` (``csrc/norm_kernels.cu``): batch statistics + apply + ReLU, and ReLU mask + both BN
  backward reductions + apply, one launch each; the Scaler is folded into epsilon;
* the conv bias in front of a train-mode BatchNorm only shifts the batch mean, which the norm subtracts again: the
  forward skips it and its gradient is identically zero (autograd produces round-off noise there);
* ``maxpool2_fwd / maxpool2_bwd`` (``csrc/pool_kernels.cu``): NHWC bf16, arg-max recomputed in the backward;
* ``head_fwd_bwd`` (``csrc/loss_kernels.cu``): average pool + linear + masked CE forward and backward;
* convolutions through the same ``ops.fused.conv_fwd / conv_bwd`` seams as the ResNet (tcgen05 wgrad where the shape
  qualifies, library otherwise); weight gradients are consumed in place by the multi-tensor optimiser.

~45 launches per step instead of ~150 through autograd.  Statistics / evaluation passes of this model run the
functional graph (``models/vision.py:conv_forward``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from ..ops import cuda as cu
from ..ops import fused as fo
from ..ops.nn import BN_EPS
from ..plan.spec import ModelSpec
from .fused_resnet import FusedResNet


class FusedConvNet:
    def __init__(self, spec: ModelSpec, scale_rate: float):
        assert FusedConvNet.eligible(spec)
        self.spec = spec
        self.nb = len(spec.meta["hidden"])
        self.scale_rate = float(scale_rate)
        self._ws: Dict[str, torch.Tensor] = {}
        self._zero: Dict[str, torch.Tensor] = {}

    @staticmethod
    def eligible(spec: ModelSpec) -> bool:
        return spec.family == "conv" and spec.meta.get("norm") == "bn"

    _scratch = FusedResNet._scratch

    @torch.no_grad()
    def train_step(self, P: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor,
                   label_mask: Optional[torch.Tensor], meter=None,
                   nvalid: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """One forward+backward; ``(loss, grads)``.  ``nvalid``: see :meth:`FusedResNet.train_step`."""
        eps = BN_EPS * self.scale_rate * self.scale_rate
        G: Dict[str, torch.Tensor] = {}
        saved: List = []
        cur = x
        for i in range(self.nb):
            cw, nm = "blocks.%d" % (5 * i), "blocks.%d" % (5 * i + 2)
            c = fo.conv_fwd(cur, P[cw + ".weight"], 1, 1)
            _, o, m, iv, _ = fo.bn_act_fwd(c, None, P[nm + ".weight"], P[nm + ".bias"], eps,
                                           ws=self._scratch(nm + ".f", c), nvalid=nvalid)
            saved.append((cw, nm, cur, c, o, m, iv))
            cur = fo.maxpool2_fwd(o) if i < self.nb - 1 else o
        N, C, H, W = cur.shape
        lw, lb = "blocks.%d.weight" % (5 * self.nb + 1), "blocks.%d.bias" % (5 * self.nb + 1)
        Wl, bl = P[lw], P[lb]
        if cur.is_cuda and cur.dtype == torch.bfloat16 and cu.available() \
                and cu.C().head_smem_bytes(N, C, Wl.shape[0]) <= 200 * 1024:
            mt, nw = (meter if meter is not None else (None, float(N)))
            mask = label_mask if (label_mask is None or label_mask.dtype == torch.uint8) else label_mask.to(torch.uint8)
            loss, G[lw], G[lb], dpooled = cu.C().head_fwd_bwd(cur, Wl, bl, y.contiguous(), mask, float(nw), mt, nvalid)
            loss = loss.view(())
        else:
            pooled = cur.mean(dim=(2, 3))
            logits = F.linear(pooled, Wl, bl)
            nvi = fo._nv(nvalid, N)
            loss, dl = FusedResNet._ce(logits[:nvi], y[:nvi], label_mask, meter)
            dlogits = torch.zeros_like(logits)
            dlogits[:nvi] = dl
            G[lw] = dlogits.t().mm(pooled)
            G[lb] = dlogits.sum(0)
            dpooled = dlogits.mm(Wl)
        g, hw = dpooled, H * W
        for i in reversed(range(self.nb)):
            cw, nm, xin, c, o, m, iv = saved[i]
            if i < self.nb - 1:
                g = fo.maxpool2_bwd(g, o)
            G[nm + ".weight"] = torch.empty_like(P[nm + ".weight"])
            G[nm + ".bias"] = torch.empty_like(P[nm + ".bias"])
            dc = fo.bn_act_bwd(g, None, None, c, m, iv, P[nm + ".weight"], P[nm + ".bias"], G[nm + ".weight"],
                               G[nm + ".bias"], hw=hw, dy_scale=(1.0 / hw) if hw else 1.0,
                               ws=self._scratch(nm + ".b", c) if hw == 0 else None, nvalid=nvalid)
            hw = 0
            g, G[cw + ".weight"] = fo.conv_bwd(dc, xin, P[cw + ".weight"], 1, 1, need_dx=i > 0)
            zb = self._zero.get(cw)                                # exactly zero behind a train-mode BatchNorm
            if zb is None or zb.shape != P[cw + ".bias"].shape or zb.device != x.device or zb.dtype != P[cw + ".bias"].dtype:
                zb = self._zero[cw] = torch.zeros_like(P[cw + ".bias"])
            G[cw + ".bias"] = zb
        return loss, G
