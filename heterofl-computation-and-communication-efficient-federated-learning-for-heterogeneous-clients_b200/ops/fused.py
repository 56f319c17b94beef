"""Fused building blocks of the manual-backward training path (GPU: ``csrc/norm_kernels.cu``; CPU: torch oracle).

Activations are logical ``[N, C, H, W]`` tensors with channels-last strides (physically ``[N*H*W, C]``).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import cuda as cu


def _use_kernels(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and cu.available()


def _f(t: torch.Tensor) -> torch.Tensor:
    """working precision of the torch oracle paths: fp32 for 16-bit inputs, unchanged otherwise (fp64 stays fp64)."""
    return t.float() if t.dtype in (torch.bfloat16, torch.float16) else t


def _rows(x: torch.Tensor) -> torch.Tensor:
    """physical [M, C] view of a channels-last activation (or of a plain [N, C] matrix)."""
    if x.dim() == 4:
        return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])
    return x.reshape(-1, x.shape[-1])


def _nv(nvalid, N: int) -> int:
    return N if nvalid is None else max(min(int(nvalid), N), 1)


def bn_act_fwd(x: torch.Tensor, res: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, eps: float,
               groups: int = 1, relu: bool = True, ws: Optional[torch.Tensor] = None, nvalid=None):
    """``s = x (+ res)``; per-group batch statistics; ``y = relu(bn(s))``.

    ``nvalid`` (int32 ``[1]`` device tensor on the kernel path, int / tensor on the oracle path): only the first
    ``nvalid`` samples of the batch are real -- the partial last batch of an epoch padded to the captured batch size.
    Statistics ignore the padding samples and ``s`` / ``y`` are zero there.

    Returns ``(s, y, mean[G,C], invstd[G,C], var_unbiased[G,C])`` (``s is x`` when ``res`` is None).
    """
    C = x.shape[1]
    if _use_kernels(x):
        s, y, mean, invstd, varu = cu.C().bn_act_fwd(x, res, gamma, beta, float(eps), int(groups), int(C), bool(relu), ws,
                                                     nvalid, int(x.shape[0]))
        return (s if res is not None else x), y, mean, invstd, varu
    s = x if res is None else (x + res)
    N = s.shape[0]
    if nvalid is not None and _nv(nvalid, N) < N:
        assert groups == 1
        nv = _nv(nvalid, N)
        sv, yv, mean, invstd, varu = bn_act_fwd(x[:nv], None if res is None else res[:nv], gamma, beta, eps, 1, relu)
        pad = lambda t: torch.cat([t, torch.zeros((N - nv,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)])
        return (pad(sv) if res is not None else x), pad(yv), mean, invstd, varu
    sf = _f(s).reshape(groups, N // groups, C, -1)
    var, mean = torch.var_mean(sf, dim=(1, 3), unbiased=False, keepdim=True)
    invstd = torch.rsqrt(var + eps)
    y = (sf - mean) * invstd * _f(gamma).view(1, 1, C, 1) + _f(beta).view(1, 1, C, 1)
    if relu:
        y = F.relu(y)
    y = y.reshape(s.shape).to(s.dtype)
    if s.dim() == 4 and s.is_cuda:
        y = y.contiguous(memory_format=torch.channels_last)
    n = sf.shape[1] * sf.shape[3]
    return s, y, mean.reshape(groups, C), invstd.reshape(groups, C), (var * (n / max(n - 1, 1))).reshape(groups, C)


def affine_act(x: torch.Tensor, res: Optional[torch.Tensor], a: torch.Tensor, b: torch.Tensor, relu: bool = True):
    """Eval-mode norm: ``s = x (+ res)``, ``y = relu(s * a[c] + b[c])``."""
    C = x.shape[1]
    if _use_kernels(x):
        s, y = cu.C().affine_act(x, res, a, b, int(C), bool(relu))
        return (s if res is not None else x), y
    s = x if res is None else (x + res)
    shape = (1, C) + (1,) * (s.dim() - 2)
    y = _f(s) * a.view(shape) + b.view(shape)
    if relu:
        y = F.relu(y)
    return s, y.to(s.dtype)


def bn_act_bwd(dy: torch.Tensor, dy2: Optional[torch.Tensor], dres: Optional[torch.Tensor], s: torch.Tensor,
               mean: torch.Tensor, invstd: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               dgamma: torch.Tensor, dbeta: torch.Tensor, hw: int = 0, dy_scale: float = 1.0,
               relu: bool = True, ws: Optional[torch.Tensor] = None, nvalid=None) -> torch.Tensor:
    """Backward of ``y = relu(bn(s))`` for one statistics group.  ``dy`` is the gradient w.r.t. ``y`` (``hw > 0``:
    ``dy`` is ``[N, C]`` and is broadcast over the ``hw`` positions of every sample, scaled by ``dy_scale``);
    ``dy2`` is an optional second gradient w.r.t. ``y``; ``dres`` an optional gradient w.r.t. ``s`` arriving through
    the identity path.  Writes ``dgamma``/``dbeta`` in place, returns ``ds``."""
    C = s.shape[1]
    if _use_kernels(s):
        return cu.C().bn_act_bwd(dy, dy2, dres, s, mean.reshape(-1), invstd.reshape(-1), gamma, beta, dgamma, dbeta,
                                 int(C), int(hw), float(dy_scale), bool(relu), ws, nvalid, int(s.shape[0]))
    if nvalid is not None and _nv(nvalid, s.shape[0]) < s.shape[0]:
        nv, N = _nv(nvalid, s.shape[0]), s.shape[0]
        cut = lambda t: None if t is None else t[:nv]
        dsv = bn_act_bwd(dy[:nv], cut(dy2), cut(dres), s[:nv], mean, invstd, gamma, beta, dgamma, dbeta, hw, dy_scale,
                         relu)
        return torch.cat([dsv, torch.zeros((N - nv,) + tuple(dsv.shape[1:]), dtype=dsv.dtype, device=dsv.device)])
    shape = (1, C) + (1,) * (s.dim() - 2)
    sf = _f(s)
    m, iv = mean.reshape(shape).float(), invstd.reshape(shape).float()
    ga, be = _f(gamma).view(shape), _f(beta).view(shape)
    xh = (sf - m) * iv
    if hw > 0:
        g = (_f(dy) * dy_scale).view(dy.shape[0], C, *([1] * (s.dim() - 2))).expand_as(sf)
    else:
        g = _f(dy) if dy2 is None else _f(dy) + _f(dy2)
    if relu:
        g = g * ((xh * ga + be) > 0)
    red = [d for d in range(s.dim()) if d != 1]
    sb, sg = g.sum(red), (g * xh).sum(red)
    n = sf.numel() // C
    ds = ga * iv * (g - sb.view(shape) / n - xh * sg.view(shape) / n)
    if dres is not None:
        ds = ds + _f(dres)
    dgamma.copy_(sg.to(dgamma.dtype))
    dbeta.copy_(sb.to(dbeta.dtype))
    ds = ds.to(s.dtype)
    if ds.dim() == 4 and ds.is_cuda:
        ds = ds.contiguous(memory_format=torch.channels_last)
    return ds


NORM_GROUPS = {"gn": lambda C: 4, "in": lambda C: C, "ln": lambda C: 1}        # same table as ops.nn.norm_act


def gn_act_fwd(x: torch.Tensor, res: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, eps: float,
               G: int, relu: bool = True):
    """``s = x (+ res)``; per-(sample, channel-group) statistics; ``y = relu(group_norm(s))`` (``csrc/gn_kernels.cu``).

    Reference: ``nn.GroupNorm`` behind the Scaler, followed by ReLU (``src/models/resnet.py:18-26``).
    Returns ``(s, y, mean[N,G], rstd[N,G])`` (``s is x`` when ``res`` is None)."""
    N, C = x.shape[0], x.shape[1]
    if _use_kernels(x):
        s, y, mean, rstd = cu.C().gn_act_fwd(x, res, gamma, beta, float(eps), int(N), int(C), int(G), bool(relu))
        return (s if res is not None else x), y, mean, rstd
    s = x if res is None else (x + res)
    sf = _f(s).reshape(N, G, -1)
    var, mean = torch.var_mean(sf, dim=2, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    shape = (1, C) + (1,) * (s.dim() - 2)
    y = ((sf - mean) * rstd).reshape(s.shape) * _f(gamma).view(shape) + _f(beta).view(shape)
    if relu:
        y = F.relu(y)
    y = y.to(s.dtype)
    if s.dim() == 4 and s.is_cuda:
        y = y.contiguous(memory_format=torch.channels_last)
    return s, y, mean.reshape(N, G), rstd.reshape(N, G)


def gn_act_bwd(dy: torch.Tensor, dy2: Optional[torch.Tensor], dres: Optional[torch.Tensor], s: torch.Tensor,
               mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               dgamma: torch.Tensor, dbeta: torch.Tensor, G: int, hw: int = 0, dy_scale: float = 1.0,
               relu: bool = True) -> torch.Tensor:
    """Backward of ``y = relu(group_norm(s))``; same gradient-join conventions as :func:`bn_act_bwd`."""
    N, C = s.shape[0], s.shape[1]
    if _use_kernels(s):
        return cu.C().gn_act_bwd(dy, dy2, dres, s, mean.contiguous(), rstd.contiguous(), gamma, beta, dgamma, dbeta,
                                 int(N), int(C), int(G), int(hw), float(dy_scale), bool(relu))
    shape = (1, C) + (1,) * (s.dim() - 2)
    sf = _f(s)
    xh = ((sf.reshape(N, G, -1) - mean.reshape(N, G, 1).to(sf.dtype)) * rstd.reshape(N, G, 1).to(sf.dtype))
    xh = xh.reshape(s.shape)
    ga, be = _f(gamma).view(shape), _f(beta).view(shape)
    if hw > 0:
        g = (_f(dy) * dy_scale).view(dy.shape[0], C, *([1] * (s.dim() - 2))).expand_as(sf)
    else:
        g = _f(dy) if dy2 is None else _f(dy) + _f(dy2)
    if relu:
        g = g * ((xh * ga + be) > 0)
    red = [d for d in range(s.dim()) if d != 1]
    dgamma.copy_((g * xh).sum(red).to(dgamma.dtype))
    dbeta.copy_(g.sum(red).to(dbeta.dtype))
    dxh = (g * ga).reshape(N, G, -1)
    xg = xh.reshape(N, G, -1)
    ds = rstd.reshape(N, G, 1).to(sf.dtype) * (dxh - dxh.mean(2, keepdim=True) - xg * (dxh * xg).mean(2, keepdim=True))
    ds = ds.reshape(s.shape)
    if dres is not None:
        ds = ds + _f(dres)
    ds = ds.to(s.dtype)
    if ds.dim() == 4 and ds.is_cuda:
        ds = ds.contiguous(memory_format=torch.channels_last)
    return ds


def maxpool2_fwd(x: torch.Tensor) -> torch.Tensor:
    """``F.max_pool2d(x, 2)`` on NHWC bf16 (``csrc/pool_kernels.cu``); reference ``src/models/conv.py:29-45``."""
    if _use_kernels(x) and x.is_contiguous(memory_format=torch.channels_last):
        return cu.C().maxpool2_fwd(x)
    return F.max_pool2d(x, 2)


def maxpool2_bwd(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Gradient of :func:`maxpool2_fwd` w.r.t. its input; the arg-max is recomputed from ``x`` (first maximum wins)."""
    if _use_kernels(x) and x.is_contiguous(memory_format=torch.channels_last) \
            and dy.is_contiguous(memory_format=torch.channels_last):
        return cu.C().maxpool2_bwd(dy, x)
    N, C, H, W = x.shape
    _, idx = F.max_pool2d(_f(x), 2, return_indices=True)
    dx = torch.zeros(N, C, H * W, dtype=dy.dtype, device=dy.device)
    dx.scatter_(2, idx.reshape(N, C, -1), dy.reshape(N, C, -1))
    return dx.reshape(N, C, H, W)


# tcgen05 implicit-GEMM forward convolution: "off" | "on" (whenever the shape qualifies) | "auto" = only where it
# measurably beats the library (profiles/conv_bench.md): the latency-bound training shapes with <= 9 K-blocks
# (C_in = 64, batch-10 tiles) | "train" = every qualifying shape of a training-sized batch (<= CONV_TC_MAX_PIXELS
# output pixels), any channel count: keeps the train step on our own kernels so that consecutive launches chain
# through programmatic dependent launch; the large-batch sBN / test passes stay on the library.
CONV_TC_MODE = os.environ.get("HFL_CONV_TC", "train")
CONV_TC_MAX_PIXELS = int(os.environ.get("HFL_CONV_TC_MAX_PIXELS", "20480"))


def _tc_conv_ok(x: torch.Tensor, w: torch.Tensor, stride: int, pad: int) -> bool:
    if CONV_TC_MODE == "off" or not _use_kernels(x) or w.dtype != torch.bfloat16:
        return False
    N, C, H, W = x.shape
    K, _, R, S = w.shape
    if pad != R // 2 or not x.is_contiguous(memory_format=torch.channels_last):
        return False
    if not w.permute(0, 2, 3, 1).is_contiguous():
        return False
    if not cu.C().conv_tc_supported(N, H, W, C, K, R, S, stride):
        return False
    if CONV_TC_MODE == "auto":
        return N * (H // stride) * (W // stride) <= CONV_TC_MAX_PIXELS and C <= 64
    if CONV_TC_MODE == "train":
        return N * (H // stride) * (W // stride) <= CONV_TC_MAX_PIXELS
    return True


def conv_fwd(x: torch.Tensor, w: torch.Tensor, stride: int, pad: int) -> torch.Tensor:
    if _tc_conv_ok(x, w, stride, pad):
        return cu.C().conv_fprop_tc(x, w, None, None, stride)
    return F.conv2d(x, w, None, stride, pad)


CONV_STATS_FUSE = os.environ.get("HFL_CONV_STATS", "1") != "0"
CONV_SC_FOLD = os.environ.get("HFL_CONV_SC_FOLD", "1") != "0"


def conv_bn_act_fwd(x: torch.Tensor, w: torch.Tensor, stride: int, pad: int, res: Optional[torch.Tensor],
                    gamma: torch.Tensor, beta: torch.Tensor, eps: float, ws: Optional[torch.Tensor], nvalid=None,
                    relu: bool = True, shortcut=None):
    """``s = conv(x, w) (+ res)``; ``y = relu(bn(s))`` with batch statistics.  Returns ``(s, y, mean, invstd)``.

    When the convolution runs on the tcgen05 kernel, its epilogue adds the residual, writes ``s`` and accumulates the
    per-channel (sum, sum of squares) of the bf16-rounded ``s`` into ``ws`` (the BN call site's persistent fp32
    ``[2C+1]`` scratch); the norm is then a single apply pass (``bn_apply_sums``) -- no statistics kernel, no separate
    residual add (reference block: ``src/models/resnet.py:44-50``).  Otherwise: library convolution + ``bn_act_fwd``.
    Padding samples (``nvalid``): the epilogue writes zeros there and keeps them out of the sums, whatever ``x`` and
    ``res`` hold in those rows.

    ``shortcut = (x_in, w_sc, stride_sc)`` instead of ``res``: the residual is the 1x1 shortcut convolution of the block
    input; on the tcgen05 path it is FOLDED into the same launch (extra k-blocks of the same accumulator tile), so a
    transition block's ``conv2(..) + shortcut(x)`` is one kernel and the shortcut output never exists in memory."""
    fused = CONV_STATS_FUSE and ws is not None and _tc_conv_ok(x, w, stride, pad) and w.shape[0] >= 8
    if shortcut is not None:
        xs, wsc, ssc = shortcut
        fold = (fused and CONV_SC_FOLD and xs.shape[1] % 8 == 0 and ssc in (1, 2)
                and xs.is_contiguous(memory_format=torch.channels_last) and wsc.permute(0, 2, 3, 1).is_contiguous()
                and xs.shape[2] == ssc * ((x.shape[2] + 2 * pad - w.shape[2]) // stride + 1))
        if fold:
            C_ = cu.C()
            K = w.shape[0]
            s = C_.conv_fprop_tc(x, w, None, ws, stride, nvalid, xs, wsc, int(ssc))
            y, mean, invstd, _ = C_.bn_apply_sums(s, gamma, beta, float(eps), int(K), bool(relu), ws, nvalid,
                                                  int(x.shape[0]))
            return s, y, mean, invstd
        res = conv_fwd(xs, wsc, ssc, 0)
    if fused:
        C_ = cu.C()
        K = w.shape[0]
        s = C_.conv_fprop_tc(x, w, res, ws, stride, nvalid)
        y, mean, invstd, _ = C_.bn_apply_sums(s, gamma, beta, float(eps), int(K), bool(relu), ws, nvalid, int(x.shape[0]))
        return s, y, mean, invstd
    c = conv_fwd(x, w, stride, pad)
    s, y, mean, invstd, _ = bn_act_fwd(c, res, gamma, beta, eps, ws=ws, nvalid=nvalid, relu=relu)
    return s, y, mean, invstd


# library convolutions: let cuDNN time its candidate kernels per shape (the reference does the same,
# src/train_classifier_fed.py:19).  The heuristic picks are poor for the narrow levels (C = 4..32: e.g. a 41 us
# wgrad2d_grouped_direct_kernel); the search runs in the warm-up steps that precede every graph capture.
# On an 8-rank box the search (8 processes timing ~400 conv configurations each while they also capture graphs) took
# 180 s of start-up against 5 s on one GPU, and the 8-GPU round is bound by the widest client's sequential chain where
# the tuned kernels gain little (252 vs 271 ms): default off from 8 ranks up, `HFL_CUDNN_BENCHMARK=1` forces it.
_default_bench = "0" if int(os.environ.get("WORLD_SIZE", "1") or 1) >= 8 else "1"
if os.environ.get("HFL_CUDNN_BENCHMARK", _default_bench) != "0":
    torch.backends.cudnn.benchmark = True

# backward convolutions on the tcgen05 kernels: "off" | "on" (dgrad + wgrad) | "wgrad" (weight gradient only)
CONV_BWD_TC_MODE = os.environ.get("HFL_CONV_BWD_TC", "on")


def _tc_bwd_ok(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, stride: int, pad: int) -> bool:
    if CONV_BWD_TC_MODE == "off" or not _use_kernels(x) or w.dtype != torch.bfloat16 or pad != w.shape[2] // 2:
        return False
    return (x.is_contiguous(memory_format=torch.channels_last) and dy.is_contiguous(memory_format=torch.channels_last)
            and w.permute(0, 2, 3, 1).is_contiguous())


def _tc_dgrad_ok(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, stride: int) -> bool:
    N, C, H, W = x.shape
    K, _, R, S = w.shape
    return CONV_BWD_TC_MODE == "on" and stride in (1, 2) and cu.C().conv_dgrad_tc_supported(N, H, W, C, K, R, S, stride)


def conv_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, stride: int, pad: int, need_dx: bool = True,
             need_dw: bool = True):
    """``(dx | None, dw)``.  tcgen05 path (``HFL_CONV_BWD_TC``): dgrad as an implicit GEMM with mirrored taps + MN-major
    filter operand (stride 2: the four output-parity classes of the strided gradient inside one launch), wgrad with
    both operands MN-major and split-K; ``dw`` then comes back as fp32 (the multi-tensor optimiser reads either
    precision).  Library path (cuDNN) for whatever the kernels do not cover."""
    if not need_dw:
        if _tc_bwd_ok(dy, x, w, stride, pad) and _tc_dgrad_ok(dy, x, w, stride):
            return cu.C().conv_dgrad_tc(dy, w, None, stride), None
        dx, _, _ = torch.ops.aten.convolution_backward(dy, x, w, None, (stride, stride), (pad, pad), (1, 1), False,
                                                       (0, 0), 1, (True, False, False))
        return dx, None
    if _tc_bwd_ok(dy, x, w, stride, pad):
        N, C, H, W = x.shape
        K, _, R, S = w.shape
        C_ = cu.C()
        dx = dw = None
        if need_dx and _tc_dgrad_ok(dy, x, w, stride):
            dx = C_.conv_dgrad_tc(dy, w, None, stride)
        if C_.conv_wgrad_tc_supported(N, H, W, C, K, R, S, stride):
            dw = C_.conv_wgrad_tc(dy, x, R, stride)
        if (dx is not None or not need_dx) and dw is not None:
            return dx, dw
        rest = torch.ops.aten.convolution_backward(dy, x, w, None, (stride, stride), (pad, pad), (1, 1), False, (0, 0), 1,
                                                   (need_dx and dx is None, dw is None, False))
        return (dx if dx is not None else rest[0]), (dw if dw is not None else rest[1])
    dx, dw, _ = torch.ops.aten.convolution_backward(dy, x, w, None, (stride, stride), (pad, pad), (1, 1), False,
                                                    (0, 0), 1, (need_dx, True, False))
    return dx, dw
