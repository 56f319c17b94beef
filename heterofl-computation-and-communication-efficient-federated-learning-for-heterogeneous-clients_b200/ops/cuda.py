"""Loader + thin Python wrappers for the in-tree sm_100a extension (``ops/_hfl_C.so``).

On a machine with a GPU the extension is mandatory: a missing/unloadable ``.so`` raises instead of silently falling
back to eager PyTorch (set ``HFL_ALLOW_FALLBACK=1`` to override for debugging).  On CPU-only machines
``available()`` is False and callers use their pure-torch paths.
"""
from __future__ import annotations

import importlib.util
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hfl_C.so")
_mod = None
_tried = False


def load():
    global _mod, _tried
    if _mod is not None or _tried:
        return _mod
    _tried = True
    if not os.path.exists(_SO):
        if torch.cuda.is_available() and os.environ.get("HFL_ALLOW_FALLBACK") != "1":
            raise ImportError("hfl_b200 CUDA extension %s is missing on a GPU machine; run "
                              "`python -m hfl_b200.ops.build` (or __graft_entry__.build())" % _SO)
        return None
    spec = importlib.util.spec_from_file_location("_hfl_C", _SO)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:
        if torch.cuda.is_available() and os.environ.get("HFL_ALLOW_FALLBACK") != "1":
            raise
        return None
    _mod = mod
    return _mod


def available() -> bool:
    """True when kernels can actually run (GPU present and extension loaded)."""
    return torch.cuda.is_available() and load() is not None


def C():
    m = load()
    if m is None:
        raise RuntimeError("hfl_b200 CUDA extension not loaded")
    return m


def launch_count() -> int:
    m = load()
    return int(m.launch_count()) if m is not None else 0


def launch_count_add(n: int) -> None:
    m = load()
    if m is not None:
        m.launch_count_add(int(n))


# -- masked cross-entropy -----------------------------------------------------------------------------------
class _MaskedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, mask, n_weight, meter):
        need = logits.requires_grad
        lg = logits.contiguous()
        loss, dl = C().masked_ce(lg, target.contiguous(), mask, float(n_weight), meter, need)
        if need:
            ctx.save_for_backward(dl)
        return loss.view(())

    @staticmethod
    def backward(ctx, gout):
        (dl,) = ctx.saved_tensors
        return dl * gout.to(dl.dtype), None, None, None, None


def masked_ce(logits: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor],
              n_weight: Optional[float] = None, meter: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused masked softmax-CE (mean over rows).  ``meter`` (float64[4] device tensor) is bumped in the same pass
    with ``[loss*n, #correct, n, exp(loss)*n]``."""
    if mask is not None and (mask.dtype != torch.uint8 or mask.device != logits.device):
        mask = mask.to(device=logits.device, dtype=torch.uint8)
    n = float(logits.shape[0] if n_weight is None else n_weight)
    return _MaskedCE.apply(logits, target, mask, n, meter)


# -- data -----------------------------------------------------------------------------------------------------
def augment_batch(src_u8: torch.Tensor, idx: Optional[torch.Tensor], params: Optional[torch.Tensor],
                  mean: torch.Tensor, inv_std: torch.Tensor, pad: int, dtype: torch.dtype) -> torch.Tensor:
    return C().augment(src_u8, idx, params, mean.reshape(-1).contiguous(), inv_std.reshape(-1).contiguous(), int(pad),
                       dtype == torch.bfloat16)


# -- optimiser ------------------------------------------------------------------------------------------------
def sgd_clip_step(w, wc, g, mom, lr_t, momentum: float, weight_decay: float, max_norm: float, scratch) -> None:
    C().sgd_clip_step(w, wc, g, mom, lr_t, float(momentum), float(weight_decay), float(max_norm), scratch)
