"""Pre-activation ResNet train / statistics / eval passes with a hand-scheduled backward.

Same network as :func:`hfl_b200.models.vision.resnet_forward` (reference ``src/models/resnet.py:140-158``) but the
step is written as an explicit sequence of fused ops instead of an autograd graph:

* ``bn_act_fwd`` = previous block's residual add + BatchNorm statistics + apply + ReLU in ONE launch
  (``csrc/norm_kernels.cu``); ``bn_act_bwd`` = gradient join + ReLU mask + BN backward reductions + apply in ONE;
* convolutions go through the library (cuDNN) for now -- ``ops.fused.conv_fwd / conv_bwd`` are the seams where the
  tcgen05 implicit-GEMM kernels plug in;
* weight gradients are consumed where their producer leaves them (segment table of the multi-tensor optimiser), so
  there is no AccumulateGrad add, no gradient copy and no zero-fill pass.

Kernel launches per train step drop from ~300 (autograd path) to ~130.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from ..ops import cuda as cu
from ..ops import fused as fo
from ..ops.nn import BN_EPS, NormCtx
from ..plan.spec import ModelSpec


class FusedResNet:
    def __init__(self, spec: ModelSpec, scale_rate: float, overlap_wgrad: bool = True):
        assert FusedResNet.eligible(spec)
        self.spec, self.blocks = spec, spec.meta["blocks"]
        self.norm = spec.meta["norm"]                       # "bn": batch statistics; "gn" / "in" / "ln": per sample
        self.scale_rate = float(scale_rate)
        self.overlap_wgrad = overlap_wgrad
        self._side = None
        self._ws: Dict[str, torch.Tensor] = {}

    def _scratch(self, name: str, like: torch.Tensor, channels: Optional[int] = None) -> Optional[torch.Tensor]:
        """Persistent, self-cleaning fp32 [2C+1] scratch of one BN call site (forward and backward use separate
        slabs); lets the wide statistics kernels run without a memset node per launch."""
        if not like.is_cuda:
            return None
        t = self._ws.get(name)
        if t is None:
            C = like.shape[1] if channels is None else int(channels)
            t = torch.zeros(2 * C + 1, dtype=torch.float32, device=like.device)
            self._ws[name] = t
        return t

    def _conv_bwd(self, dy, x, w, stride, pad, keep, need_dx=True):
        """dgrad on the current stream (it feeds the next layer's backward: the critical path), wgrad on a side stream
        (only the optimiser needs it).  ``keep`` holds every tensor the side stream reads until the step's join."""
        if not (self.overlap_wgrad and dy.is_cuda):
            return fo.conv_bwd(dy, x, w, stride, pad, need_dx=need_dx)
        main = torch.cuda.current_stream(dy.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dy.device)
        keep.extend((dy, x))
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            _, dw = fo.conv_bwd(dy, x, w, stride, pad, need_dx=False)
        dx = fo.conv_bwd(dy, x, w, stride, pad, need_dx=True, need_dw=False)[0] if need_dx else None
        return dx, dw

    @staticmethod
    def eligible(spec: ModelSpec, train_only: bool = True) -> bool:
        """BatchNorm models use the fused path for training, the sBN pass and evaluation; GroupNorm / InstanceNorm /
        LayerNorm models for the training step (their evaluation has no statistics pass and runs the functional
        model)."""
        norms = ("bn", "gn", "in", "ln") if train_only else ("bn",)
        return spec.family == "resnet" and spec.meta.get("norm") in norms and spec.meta.get("kind") == "block"

    def _nf(self, x, res, P, name, eps, ws_tag, nvalid=None):
        """fused (residual add +) norm + ReLU forward -> (s, y, saved statistics)"""
        if self.norm == "bn":
            s, y, m, i, _ = fo.bn_act_fwd(x, res, P[name + ".weight"], P[name + ".bias"], eps,
                                          ws=self._scratch(ws_tag, x), nvalid=nvalid)
        else:
            s, y, m, i = fo.gn_act_fwd(x, res, P[name + ".weight"], P[name + ".bias"], eps,
                                       fo.NORM_GROUPS[self.norm](x.shape[1]))
        return s, y, m, i

    def _cnf(self, pend, P, name, eps, ws_tag, nvalid=None):
        """3x3 convolution ``pend = (input, weight, stride, residual)`` followed by ``_nf`` -> (s, y, saved stats);
        BatchNorm models fuse the residual add and the batch statistics into the convolution's epilogue."""
        xin, w, stride, res = pend
        shortcut = None
        if isinstance(res, tuple):                        # ("conv", block input, shortcut filter, stride)
            shortcut, res = res[1:], None
        if self.norm == "bn":
            return fo.conv_bn_act_fwd(xin, w, stride, 1, res, P[name + ".weight"], P[name + ".bias"], eps,
                                      self._scratch(ws_tag, xin, w.shape[0]), nvalid, shortcut=shortcut)
        if shortcut is not None:
            res = fo.conv_fwd(shortcut[0], shortcut[1], shortcut[2], 0)
        return self._nf(fo.conv_fwd(xin, w, stride, 1), res, P, name, eps, ws_tag, nvalid)

    def _nb(self, dy, dy2, dres, s, m, i, P, G, name, ws_tag, nvalid=None, **kw):
        G[name + ".weight"] = torch.empty_like(P[name + ".weight"])
        G[name + ".bias"] = torch.empty_like(P[name + ".bias"])
        if self.norm == "bn":
            return fo.bn_act_bwd(dy, dy2, dres, s, m, i, P[name + ".weight"], P[name + ".bias"], G[name + ".weight"],
                                 G[name + ".bias"], ws=self._scratch(ws_tag, s) if ws_tag else None, nvalid=nvalid, **kw)
        return fo.gn_act_bwd(dy, dy2, dres, s, m, i, P[name + ".weight"], P[name + ".bias"], G[name + ".weight"],
                             G[name + ".bias"], fo.NORM_GROUPS[self.norm](s.shape[1]), **kw)

    # ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def train_step(self, P: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor,
                   label_mask: Optional[torch.Tensor], meter=None,
                   nvalid: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """One forward+backward.  Returns ``(loss, grads)`` with a gradient tensor for every parameter.

        ``nvalid`` (int32 ``[1]`` on the device): the batch is padded and only its first ``nvalid`` samples are real
        (the partial last batch of an epoch running through the SAME captured graph as the full ones).  BatchNorm
        statistics, the loss mean and the meter use ``nvalid``; the padding rows are zeroed after every norm in the
        forward and the backward pass, so they contribute exactly nothing to any weight gradient.  Per-sample norms
        (GroupNorm family) need no masking beyond the head: a zero output gradient stays zero."""
        nv = nvalid
        s_rate = self.scale_rate
        eps = BN_EPS * s_rate * s_rate          # Scaler folded into epsilon
        G: Dict[str, torch.Tensor] = {}
        saved = []
        keep: List[torch.Tensor] = []
        # Every norm consumes the convolution in front of it (+ the block's shortcut): conv -> (residual add + batch
        # statistics in the tcgen05 epilogue) -> apply + ReLU.  `pend` = (input, weight, residual) of the convolution
        # whose output the NEXT norm normalises.
        c0 = fo.conv_fwd(x, P["conv1.weight"], 1, 1)
        s_in = o1 = m1 = i1 = None
        pend = None
        for bi, blk in enumerate(self.blocks):
            pre, stride = blk["prefix"], blk["stride"]
            if pend is None:                    # first block: normalises the stem convolution
                s_in, o1, m1, i1 = self._nf(c0, None, P, pre + ".n1", eps, pre + ".n1.f", nv)
            else:
                s_in, o1, m1, i1 = self._cnf(pend, P, pre + ".n1", eps, pre + ".n1.f", nv)
            # residual of this block's output: the block input, or (transition blocks) its 1x1 shortcut convolution --
            # handed to the NEXT norm's convolution as an operand pair so that it is folded into the same launch
            sc = ("conv", o1, P[pre + ".shortcut.weight"], stride) if blk["shortcut"] else s_in
            c1, o2, m2, i2 = self._cnf((o1, P[pre + ".conv1.weight"], stride, None), P, pre + ".n2", eps, pre + ".n2.f",
                                       nv)
            saved.append((blk, s_in, o1, m1, i1, c1, o2, m2, i2))
            pend = (o2, P[pre + ".conv2.weight"], 1, sc)
        s4, y4, m4, i4 = self._cnf(pend, P, "n4", eps, "n4.f", nv)
        # ---- head: global average pool -> linear -> masked CE (forward + backward) ----
        N, C, H, W = y4.shape
        Wl, bl = P["linear.weight"], P["linear.bias"]
        if y4.is_cuda and y4.dtype == torch.bfloat16 and cu.available() \
                and cu.C().head_smem_bytes(N, C, Wl.shape[0]) <= 200 * 1024:
            mt, nw = (meter if meter is not None else (None, float(N)))
            mask = label_mask if (label_mask is None or label_mask.dtype == torch.uint8) else label_mask.to(torch.uint8)
            loss, G["linear.weight"], G["linear.bias"], dpooled = cu.C().head_fwd_bwd(
                y4, Wl, bl, y.contiguous(), mask, float(nw), mt, nv)
            loss = loss.view(())
        else:
            pooled = y4.mean(dim=(2, 3))
            logits = F.linear(pooled, Wl, bl)
            nvi = fo._nv(nv, N)
            loss, dl = self._ce(logits[:nvi], y[:nvi], label_mask, meter)
            dlogits = torch.zeros_like(logits)
            dlogits[:nvi] = dl
            G["linear.weight"] = dlogits.t().mm(pooled)
            G["linear.bias"] = dlogits.sum(0)
            dpooled = dlogits.mm(Wl)                                  # [N, C]
        # ---- backward ----
        ds = self._nb(dpooled, None, None, s4, m4, i4, P, G, "n4", None, nv, hw=H * W, dy_scale=1.0 / (H * W))
        for (blk, s_in, o1, m1, i1, c1, o2, m2, i2) in reversed(saved):
            pre, stride = blk["prefix"], blk["stride"]
            # ds = grad wrt block output (= c2 + sc)
            do2, G[pre + ".conv2.weight"] = self._conv_bwd(ds, o2, P[pre + ".conv2.weight"], 1, 1, keep)
            dc1 = self._nb(do2, None, None, c1, m2, i2, P, G, pre + ".n2", pre + ".n2.b", nv)
            do1, G[pre + ".conv1.weight"] = self._conv_bwd(dc1, o1, P[pre + ".conv1.weight"], stride, 1, keep)
            if blk["shortcut"]:
                do1b, G[pre + ".shortcut.weight"] = self._conv_bwd(ds, o1, P[pre + ".shortcut.weight"], stride, 0, keep)
                ds = self._nb(do1, do1b, None, s_in, m1, i1, P, G, pre + ".n1", pre + ".n1.b", nv)
            else:
                ds = self._nb(do1, None, ds, s_in, m1, i1, P, G, pre + ".n1", pre + ".n1.b", nv)
        _, G["conv1.weight"] = self._conv_bwd(ds, x, P["conv1.weight"], 1, 1, keep, need_dx=False)
        if self._side is not None and x.is_cuda:
            torch.cuda.current_stream(x.device).wait_stream(self._side)    # join: all weight gradients are complete
        return loss, G

    @staticmethod
    def _ce(logits, y, label_mask, meter):
        if logits.is_cuda and cu.available():
            mt, nw = (meter if meter is not None else (None, float(logits.shape[0])))
            mask = label_mask
            if mask is not None and mask.dtype != torch.uint8:
                mask = mask.to(torch.uint8)
            loss, dl = cu.C().masked_ce(logits.contiguous(), y.contiguous(), mask, float(nw), mt, True)
            return loss.view(()), dl
        z = logits if logits.dtype == torch.float64 else logits.float()
        if label_mask is not None:
            keep = label_mask.to(z.device).bool().view(1, -1)
            z = z.masked_fill(~keep, 0.0)
        lse = torch.logsumexp(z, dim=1)
        loss = (lse - z.gather(1, y.view(-1, 1)).squeeze(1)).mean()
        dl = torch.softmax(z, dim=1)
        dl[torch.arange(z.shape[0]), y] -= 1.0
        dl = dl / z.shape[0]
        if label_mask is not None:
            dl = dl * keep
        return loss, dl.to(logits.dtype)

    # ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_stats(self, P: Dict[str, torch.Tensor], x: torch.Tensor, ctx: NormCtx) -> None:
        """sBN statistics pass over ``groups = N // ctx.group`` ghost batches; accumulates into ``ctx.acc``."""
        s_rate = self.scale_rate
        eps = BN_EPS * s_rate * s_rate
        Gn = max(x.shape[0] // ctx.group, 1)

        def acc(name, mean, varu):
            m, v = mean.sum(0), varu.sum(0)
            if s_rate != 1.0:
                m, v = m / s_rate, v / (s_rate * s_rate)
            if name in ctx.acc:
                ctx.acc[name][0].add_(m)
                ctx.acc[name][1].add_(v)
            else:
                ctx.acc[name] = (m.clone(), v.clone())

        cur, res = fo.conv_fwd(x, P["conv1.weight"], 1, 1), None
        for blk in self.blocks:
            pre, stride = blk["prefix"], blk["stride"]
            s_in, o1, m1, _, v1 = fo.bn_act_fwd(cur, res, P[pre + ".n1.weight"], P[pre + ".n1.bias"], eps, Gn)
            acc(pre + ".n1", m1, v1)
            sc = fo.conv_fwd(o1, P[pre + ".shortcut.weight"], stride, 0) if blk["shortcut"] else s_in
            c1 = fo.conv_fwd(o1, P[pre + ".conv1.weight"], stride, 1)
            _, o2, m2, _, v2 = fo.bn_act_fwd(c1, None, P[pre + ".n2.weight"], P[pre + ".n2.bias"], eps, Gn)
            acc(pre + ".n2", m2, v2)
            cur, res = fo.conv_fwd(o2, P[pre + ".conv2.weight"], 1, 1), sc
        _, _, m4, _, v4 = fo.bn_act_fwd(cur, res, P["n4.weight"], P["n4.bias"], eps, Gn)
        acc("n4", m4, v4)

    @staticmethod
    def eval_affine(P: Dict[str, torch.Tensor], running: Dict) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
        """Per-BN ``(a, b)`` with ``y = x*a + b`` from running statistics (computed once per round)."""
        out = {}
        for name, (rm, rv) in running.items():
            a = P[name + ".weight"].float() * torch.rsqrt(rv.float() + BN_EPS)
            out[name] = (a.contiguous(), (P[name + ".bias"].float() - rm.float() * a).contiguous())
        return out

    @torch.no_grad()
    def forward_eval(self, P: Dict[str, torch.Tensor], x: torch.Tensor, ab: Dict) -> torch.Tensor:
        """Eval-mode logits (running statistics, Scaler = identity)."""
        cur, res = fo.conv_fwd(x, P["conv1.weight"], 1, 1), None
        for blk in self.blocks:
            pre, stride = blk["prefix"], blk["stride"]
            s_in, o1 = fo.affine_act(cur, res, *ab[pre + ".n1"])
            sc = fo.conv_fwd(o1, P[pre + ".shortcut.weight"], stride, 0) if blk["shortcut"] else s_in
            c1 = fo.conv_fwd(o1, P[pre + ".conv1.weight"], stride, 1)
            _, o2 = fo.affine_act(c1, None, *ab[pre + ".n2"])
            cur, res = fo.conv_fwd(o2, P[pre + ".conv2.weight"], 1, 1), sc
        _, y4 = fo.affine_act(cur, res, *ab["n4"])
        return F.linear(y4.mean(dim=(2, 3)), P["linear.weight"], P["linear.bias"]).float()
