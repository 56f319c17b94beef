"""Masked-LM Transformer encoder as a function of a parameter dict.

Parity: reference ``src/models/transformer.py:11-174`` -- learned positional table of ``bptt`` rows, token table
with one extra [MASK] row, post-LN encoder layers with a hand-rolled MHA (separate q/k/v/o projections,
temperature ``sqrt(E // H)``), GELU FFN, decoder ``linear1 -> GELU -> LN -> linear2(vocab)``; 15 % of the input
positions are replaced by the [MASK] id *in every forward* (train and eval), attention is unmasked, the loss is CE
over ALL positions against the original tokens.  Scaler placement as in the reference, but folded: the
pre-LayerNorm scalers become an epsilon, the projection scalers become GEMM ``alpha`` factors and the q/k scalers
fold into the softmax scale.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from ..ops.linear import tc_linear
from ..ops.nn import NormCtx, layer_norm_scaled, masked_cross_entropy, masked_logits
from ..plan.spec import ModelSpec


def _linear(x, w, b, alpha: float = 1.0):
    """``alpha * (x W^T + b)`` -- tcgen05 GEMM with the Scaler as its alpha on the GPU, F.linear elsewhere."""
    return tc_linear(x, w, b, alpha)


def transformer_forward(spec: ModelSpec, P: Dict[str, torch.Tensor], inp: Dict, ctx: NormCtx, scale_rate: float,
                        label_mask: Optional[torch.Tensor] = None, want_score: bool = False,
                        mask_rate: float = 0.15, dropout: float = 0.2) -> Dict:
    m = spec.meta
    H, V, L = m["heads"], m["num_tokens"], m["layers"]
    train = ctx.training
    s = float(scale_rate) if train else 1.0
    inv = 1.0 / s
    p_drop = dropout if train else 0.0
    tokens = inp["label"]
    N, S = tokens.shape
    if "mlm_mask" in inp:
        mlm = inp["mlm_mask"]
    else:
        mlm = torch.rand((N, S), device=tokens.device) < mask_rate
    src = tokens.masked_fill(mlm, V)
    emb = P["transformer_embedding.embedding.weight"]
    pos = P["transformer_embedding.positional_embedding.positional_embedding.weight"]
    x = F.embedding(src, emb) + pos[:S].unsqueeze(0)
    x = layer_norm_scaled(x, P["transformer_embedding.norm.weight"], P["transformer_embedding.norm.bias"], s, train)
    x = F.dropout(x, p_drop, train)
    E = x.shape[-1]
    dh = E // H
    temp = math.sqrt(dh)
    for l in range(L):
        pre = "transformer_encoder.layers.%d" % l
        q = _linear(x, P[pre + ".mha.linear_q.weight"], P[pre + ".mha.linear_q.bias"])
        k = _linear(x, P[pre + ".mha.linear_k.weight"], P[pre + ".mha.linear_k.bias"])
        v = _linear(x, P[pre + ".mha.linear_v.weight"], P[pre + ".mha.linear_v.bias"], inv)
        q = q.view(N, S, H, dh).transpose(1, 2)
        k = k.view(N, S, H, dh).transpose(1, 2)
        v = v.view(N, S, H, dh).transpose(1, 2)
        # (q/s)(k/s)^T / temp: the q/k scalers fold into the softmax scale; v/s and the post-projection
        # scaler are plain alpha factors
        att = torch.softmax((q @ k.transpose(-2, -1)) * (inv * inv / temp), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(N, S, E)      # v already carries its 1/s (GEMM alpha)
        o = _linear(o, P[pre + ".mha.linear_o.weight"], P[pre + ".mha.linear_o.bias"], inv)
        x = F.layer_norm(x + F.dropout(o, p_drop, train), (E,), P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
        h = F.gelu(_linear(x, P[pre + ".linear1.weight"], P[pre + ".linear1.bias"], inv))
        h = _linear(F.dropout(h, p_drop, train), P[pre + ".linear2.weight"], P[pre + ".linear2.bias"], inv)
        x = F.layer_norm(x + F.dropout(h, p_drop, train), (E,), P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    d = F.gelu(_linear(x, P["decoder.linear1.weight"], P["decoder.linear1.bias"], inv))
    d = F.layer_norm(d, (E,), P["decoder.norm1.weight"], P["decoder.norm1.bias"])
    logits = _linear(d, P["decoder.linear2.weight"], P["decoder.linear2.bias"]).reshape(N * S, V)
    if want_score == "raw":            # evaluation: the caller reduces per window
        return {"logits": logits}
    out = {"loss": masked_cross_entropy(logits, tokens.reshape(-1), label_mask, inp.get("_meter"))}
    if want_score:
        out["score"] = masked_logits(logits.float(), label_mask).view(N, S, V).permute(0, 2, 1)
    return out
