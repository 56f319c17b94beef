"""Linear layers on the tcgen05 GEMM (``csrc/gemm_tcgen05.cu``).

``tc_linear(x, w, b, alpha)`` = ``alpha * (x @ w^T + b)`` -- the Scaler of the reference's Transformer
(``src/models/transformer.py:80-84,116,132``: ``scaler(linear(x))``) rides along as the GEMM's ``alpha`` instead of a
separate elementwise kernel.  Forward runs on the hand-written TMA/TMEM kernel whenever the operands qualify (CUDA,
bf16, K multiple of 8, 16-byte aligned rows); the backward GEMMs read their transposed operands in place through
MN-major UMMA descriptors (``gemm_tcgen05_general_kernel``).  Falls back to ``F.linear`` everywhere else (CPU, fp32,
odd shapes).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import cuda as cu

_ENABLED = True


def set_enabled(flag: bool) -> None:
    global _ENABLED
    _ENABLED = bool(flag)


def _qualifies(x2: torch.Tensor, w: torch.Tensor) -> bool:
    if not (_ENABLED and x2.is_cuda and x2.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and cu.available()):
        return False
    K = x2.shape[1]
    return (K % 8 == 0 and x2.stride(1) == 1 and w.stride(1) == 1 and x2.stride(0) % 8 == 0 and w.stride(0) % 8 == 0
            and x2.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0)


class _TCLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, w, b, alpha):
        ctx.save_for_backward(x2, w)
        ctx.alpha, ctx.has_bias = alpha, b is not None
        return cu.C().gemm_bf16_tn(x2, w, None if b is None else b * alpha if alpha != 1.0 else b, float(alpha), False,
                                   False)

    @staticmethod
    def backward(ctx, dy):
        # backward GEMMs on the same tcgen05 kernel family; the transposed operands are consumed in place through
        # MN-major UMMA descriptors (no transpose kernels):
        #   dX[M,K] = alpha * dY[M,N] . W[N,K]      (B = W read as [red=N, out=K]: MN-major)
        #   dW[N,K] = alpha * dY[M,N]^T . X[M,K]    (A = dY as [red=M, out=N], B = X as [red=M, out=K]: both MN-major)
        x2, w = ctx.saved_tensors
        a = float(ctx.alpha)
        dy = dy.contiguous()
        M, N = dy.shape
        ok = (M % 8 == 0 and N % 8 == 0 and dy.data_ptr() % 16 == 0 and w.stride(0) % 8 == 0 and x2.stride(0) % 8 == 0)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = cu.C().gemm_bf16_general(dy, False, w, True, a, False) if ok else (dy.mm(w) * a)
        if ctx.needs_input_grad[1]:
            dw = cu.C().gemm_bf16_general(dy, True, x2, True, a, False) if ok else (dy.t().mm(x2) * a)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0) if a == 1.0 else dy.sum(0) * a
        return dx, dw, db, None


def tc_linear(x: torch.Tensor, w: torch.Tensor, b=None, alpha: float = 1.0) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if _qualifies(x2, w):
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = _TCLinear.apply(x2, w, b, float(alpha))
        return y.view(*x.shape[:-1], w.shape[0])
    y = F.linear(x, w, b)
    return y if alpha == 1.0 else y * alpha
