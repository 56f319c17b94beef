"""Neural-net building blocks with the HeteroFL twists folded in.

* **Scaler folding.**  The reference divides activations by ``rate`` right before every norm layer
  (``src/modules/modules.py:9-11``; placement ``src/models/resnet.py:45,48,148``).  ``norm(x / s)`` equals
  ``norm(x)`` computed with ``eps * s**2``, so the Scaler costs nothing here: it becomes an epsilon.  Only with
  ``norm='none'`` a real multiply remains (fused into the ReLU).
* **Static BatchNorm.**  Training never tracks running stats (``momentum=None, track_running_stats=False``,
  ``src/models/resnet.py:16-17``).  The post-round statistics pass (``src/train_classifier_fed.py:127-138``)
  feeds the train set through the model in batches of ``B``; running stats become the plain average of
  per-batch means and per-batch *unbiased* variances.  :class:`NormCtx` in ``stats`` mode evaluates G such
  batches in one launch ("ghost" groups) and accumulates the sums on device.
* **Masked cross-entropy.**  Logits of classes outside the client's label set are *zero-filled* (not -inf) before
  softmax-CE (``src/models/resnet.py:152-157``).

Every function has a pure-torch path (CPU / oracle) and dispatches to the fused sm_100a kernels in
``hfl_b200.ops.cuda`` when the tensors live on a GPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


class NormCtx:
    """How normalisation layers behave for one forward pass.

    mode 'train': batch statistics (the whole batch is one group).
    mode 'stats': batch statistics per ghost group of ``group`` samples; sums of per-group mean / unbiased var
                  accumulate into ``acc[name] = [sum_mean, sum_var]`` and ``n_groups``.
    mode 'eval' : running statistics from ``running[name] = (mean, var)``; Scaler is the identity.
    """

    def __init__(self, mode: str = "train", group: int = 0, running: Optional[Dict] = None):
        assert mode in ("train", "stats", "eval")
        self.mode, self.group = mode, group
        self.running = running if running is not None else {}
        self.acc: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.n_groups = 0

    @property
    def training(self) -> bool:
        return self.mode != "eval"

    def finalize(self) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
        """running stats = mean over ghost groups (cumulative moving average with momentum=None)."""
        return {k: (m / max(self.n_groups, 1), v / max(self.n_groups, 1)) for k, (m, v) in self.acc.items()}


def _group_batch_norm(x: torch.Tensor, G: int, weight, bias, eps: float, want_stats: bool):
    """BatchNorm over ghost groups: x [G*n, C, H, W]; statistics per (group, channel)."""
    N, C = x.shape[0], x.shape[1]
    n = N // G
    x5 = x.reshape(G, n, C, -1).float()
    var, mean = torch.var_mean(x5, dim=(1, 3), unbiased=False, keepdim=True)
    y = (x5 - mean) * torch.rsqrt(var + eps)
    y = y * weight.float().view(1, 1, C, 1) + bias.float().view(1, 1, C, 1)
    y = y.reshape(x.shape).to(x.dtype)
    if x.dim() == 4 and x.is_cuda:
        y = y.contiguous(memory_format=torch.channels_last)
    if not want_stats:
        return y, None, None
    cnt = n * x5.shape[3]
    unb = var * (cnt / max(cnt - 1, 1))
    return y, mean.reshape(G, C), unb.reshape(G, C)


def norm_act(x: torch.Tensor, kind: str, weight, bias, name: str, ctx: NormCtx, scale_rate: float = 1.0,
             relu: bool = True) -> torch.Tensor:
    """``relu(norm(x / scale_rate))`` with the Scaler folded into epsilon (identity when not training)."""
    s = float(scale_rate) if ctx.training else 1.0
    if kind == "none":
        y = x if s == 1.0 else x * (1.0 / s)
        return F.relu(y) if relu else y
    if kind == "bn":
        eps = BN_EPS * s * s
        if ctx.mode == "eval":
            rm, rv = ctx.running[name]
            y = F.batch_norm(x, rm.float(), rv.float(), weight.float(), bias.float(), False, 0.0, BN_EPS)
        elif ctx.mode == "train":
            y = F.batch_norm(x, None, None, weight, bias, True, 0.0, eps)
        else:  # stats
            G = max(x.shape[0] // ctx.group, 1)
            y, m, v = _group_batch_norm(x, G, weight, bias, eps, True)
            if s != 1.0:  # statistics are those of the *scaled* activations x / s
                m, v = m / s, v / (s * s)
            if name in ctx.acc:
                ctx.acc[name][0].add_(m.sum(0))
                ctx.acc[name][1].add_(v.sum(0))
            else:
                ctx.acc[name] = (m.sum(0), v.sum(0))
        return F.relu(y) if relu else y
    C = x.shape[1]
    groups = {"in": C, "ln": 1, "gn": 4}[kind]
    y = F.group_norm(x, groups, weight, bias, BN_EPS * s * s)
    return F.relu(y) if relu else y


def layer_norm_scaled(x: torch.Tensor, weight, bias, scale_rate: float, training: bool) -> torch.Tensor:
    """``LayerNorm(x / s)`` == LayerNorm with eps*s^2."""
    s = float(scale_rate) if training else 1.0
    return F.layer_norm(x, (x.shape[-1],), weight, bias, 1e-5 * s * s)


def masked_cross_entropy(logits: torch.Tensor, target: torch.Tensor, label_mask: Optional[torch.Tensor],
                         meter=None) -> torch.Tensor:
    """Mean CE over rows of ``logits [M, V]`` after zero-filling masked-out classes.

    ``label_mask``: uint8/bool ``[V]`` (1 = class present on this client) or None.
    ``meter``: optional ``(float64[4] device tensor, n_weight)`` bumped inside the fused kernel with
    ``[loss*n, #correct, n, exp(loss)*n]`` (GPU path only).
    """
    if logits.is_cuda:
        from . import cuda as cu
        if cu.available():
            if meter is not None:
                return cu.masked_ce(logits, target, label_mask, meter[1], meter[0])
            return cu.masked_ce(logits, target, label_mask)
    z = logits if logits.dtype == torch.float64 else logits.float()
    if label_mask is not None:
        z = z.masked_fill(label_mask.to(z.device).bool().logical_not().view(1, -1), 0.0)
    return F.cross_entropy(z, target)


def masked_logits(logits: torch.Tensor, label_mask: Optional[torch.Tensor]) -> torch.Tensor:
    if label_mask is None:
        return logits
    return logits.masked_fill(label_mask.to(logits.device).bool().logical_not().view(1, -1), 0.0)
