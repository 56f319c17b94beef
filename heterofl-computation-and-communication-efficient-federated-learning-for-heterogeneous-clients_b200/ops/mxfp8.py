"""MXFP8 (OCP microscaling: e4m3 elements, one power-of-two ue8m0 scale per 32 elements along K) linear algebra.

GPU: ``csrc/gemm_mxfp8_tcgen05.cu`` (``tcgen05.mma.kind::mxf8f6f4.block_scale`` with the scale factors in TMEM).  This
module also holds the plain-PyTorch oracle of the format -- quantiser, scale-factor layout, dequantiser -- that the tests
compare the kernels against, and :func:`mx_linear`, the drop-in used by the language model when ``HFL_LM_FP8=1``.

Reference call sites: the Transformer's ``nn.Linear`` layers (``src/models/transformer.py:59-62,94-96,125-129``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import cuda as cu

E4M3_MAX = 448.0
BLOCK = 32


def sf_offsets(R: int, K: int, device=None) -> torch.Tensor:
    """Byte offset of the scale factor of (row r, 32-block kb32) in the tcgen05 layout: one 512-byte chunk per
    (128-row tile, 128-element k-block), inside it ``(r % 32) * 16 + (r % 128 // 32) * 4 + kb32 % 4``."""
    r = torch.arange(R, device=device).view(-1, 1)
    kb = torch.arange(K // BLOCK, device=device).view(1, -1)
    return ((r // 128) * (K // 128) + kb // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + kb % 4


def quantize_ref(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Oracle quantiser: ``(q float8_e4m3fn [R, K], e int32 [R, K/32])`` with ``x ~= q * 2**e``."""
    R, K = x.shape
    xb = x.float().view(R, K // BLOCK, BLOCK)
    amax = xb.abs().amax(dim=2)
    e = torch.where(amax > 0, torch.floor(torch.log2(amax.clamp_min(1e-45))) - 8, torch.full_like(amax, -127.0))
    e = e.clamp(-127, 127)
    q = (xb * torch.exp2(-e).unsqueeze(2)).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q.view(R, K), e.to(torch.int32)


def dequantize(q_u8: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    """fp32 ``[R, K]`` from the kernel-side representation (e4m3 bit patterns + laid-out ue8m0 bytes)."""
    R, K = q_u8.shape
    e = sf[sf_offsets(R, K, sf.device)].to(torch.int32) - 127
    q = q_u8.view(torch.float8_e4m3fn).float().view(R, K // BLOCK, BLOCK)
    return (q * torch.exp2(e.float()).unsqueeze(2)).view(R, K)


def quantize(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kernel-side quantiser: ``(q uint8 [R, K], sf uint8 [...])``; torch oracle in the same layout off the GPU."""
    if x.is_cuda and x.dtype == torch.bfloat16 and cu.available():
        return tuple(cu.C().mx_quantize(x.contiguous()))
    q, e = quantize_ref(x)
    R, K = x.shape
    sf = torch.zeros(((R + 127) // 128) * (K // 128) * 512, dtype=torch.uint8, device=x.device)
    sf[sf_offsets(R, K, x.device).reshape(-1)] = (e + 127).to(torch.uint8).reshape(-1)
    return q.view(torch.uint8), sf


def mx_gemm(aq, sfa, bq, sfb, bias: Optional[torch.Tensor] = None, alpha: float = 1.0, out_fp32: bool = False):
    if aq.is_cuda and cu.available():
        return cu.C().gemm_mxfp8(aq, sfa, bq, sfb, bias, float(alpha), bool(out_fp32))
    c = dequantize(aq, sfa) @ dequantize(bq, sfb).t() * alpha
    if bias is not None:
        c = c + bias
    return c if out_fp32 else c.bfloat16()


def mx_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
              wq: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """``x @ w.T + bias`` with both operands quantised to MXFP8 on the fly (``wq``: pre-quantised weight)."""
    aq, sfa = quantize(x)
    bq, sfb = wq if wq is not None else quantize(w)
    return mx_gemm(aq, sfa, bq, sfb, None if bias is None else bias.float(), 1.0, False)
