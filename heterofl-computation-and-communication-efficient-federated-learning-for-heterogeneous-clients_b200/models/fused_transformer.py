"""Masked-LM Transformer train step with a hand-scheduled backward on fused kernels.

Same network as :func:`hfl_b200.models.transformer.transformer_forward` (reference ``src/models/transformer.py:11-174``)
but written as an explicit kernel sequence instead of an autograd graph.  The federated LM step is 1 row x 64 tokens
per client -- it is bound by the NUMBER of launches (447 per step through autograd), so everything between the GEMMs is
collapsed (``csrc/lm_kernels.cu``):

=====================  ================================================================================================
``lm_embed_fwd/bwd``   in-kernel Bernoulli [MASK] replacement + token/position gather + LayerNorm(eps*s^2) + dropout;
                       backward scatter-adds into the (fp32, zero-filled) embedding gradient
``lm_attn_fwd/bwd``    softmax(q k^T * 1/(s^2 sqrt(dh))) v per (row, head) in one CTA; backward also reduces the q/k/v
                       bias gradients
``lm_add_ln_fwd/bwd``  LayerNorm(x + dropout(h)); backward joins up to four gradient branches on load and emits the
                       residual gradient, the branch gradient, dgamma/dbeta and the producing linear's bias gradient
``lm_gelu_drop_*``     dropout(GELU(u)) and its backward + bias gradient
tcgen05 GEMMs          ``alpha * (x W^T + b)`` with the bf16 bias read straight from the parameter arena; dX / dW read
                       their transposed operands in place through MN-major UMMA descriptors
``masked_ce``          zero-fill masked CE forward + dlogits + metrics (row-padded logits so V = 33 278 needs no copy)
=====================  ================================================================================================

Weight gradients stay where their producing GEMM left them (bf16); bias / LayerNorm / embedding gradients are fp32
atomics into two pooled buffers that one memset clears per step.  The multi-tensor optimiser consumes all of them in
place (``fed/trainer.py``).  ~125 launches per step instead of 447.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from ..ops import cuda as cu
from ..plan.spec import ModelSpec

EMB = "transformer_embedding.embedding.weight"
POS = "transformer_embedding.positional_embedding.positional_embedding.weight"
LN_EPS = 1e-5


def _aligned(t: torch.Tensor) -> bool:
    return t.data_ptr() % 16 == 0 and t.stride(-1) == 1 and (t.dim() < 2 or t.stride(0) % 8 == 0)


class FusedTransformer:
    is_lm = True

    @staticmethod
    def eligible(spec: ModelSpec, emb_width: int) -> bool:
        """``emb_width`` = this sub-network's embedding size (the LayerNorm kernels keep a row in registers)."""
        if spec.family != "transformer" or not cu.available():
            return False
        m = spec.meta
        return (emb_width <= cu.C().lm_max_e() and emb_width % m["heads"] == 0 and emb_width // m["heads"] <= 32
                and m["bptt"] <= 64)

    def __init__(self, spec: ModelSpec, scale_rate: float, mask_rate: float = 0.15, dropout: float = 0.2):
        assert spec.family == "transformer"
        self.spec = spec
        m = spec.meta
        self.H, self.V, self.L = int(m["heads"]), int(m["num_tokens"]), int(m["layers"])
        self.scale_rate = float(scale_rate)
        self.mask_rate, self.dropout = float(mask_rate), float(dropout)
        self._small: Optional[torch.Tensor] = None
        self._big: Optional[torch.Tensor] = None
        self._small_views: Dict[str, torch.Tensor] = {}
        self._big_views: Dict[str, torch.Tensor] = {}
        self._logits: Dict[int, torch.Tensor] = {}
        import os
        self.grouped = os.environ.get("HFL_LM_GROUPED", "1") != "0"     # grouped q/k/v GEMMs (forward, dW, dX)
        # opt-in reduced precision: forward GEMMs of the feed-forward blocks in MXFP8 (e4m3 + ue8m0 block scales,
        # csrc/gemm_mxfp8_tcgen05.cu); backward stays bf16.  Off by default: bench.py's dtype is bf16.
        self.fp8 = os.environ.get("HFL_LM_FP8", "0") == "1"
        # fused vocabulary head (csrc/lm_vocab_tcgen05.cu): decoder.linear2 + label mask + cross-entropy forward and
        # backward in three launches, no [tokens, vocabulary] tensor in memory
        self.vocab_fused = os.environ.get("HFL_LM_VOCAB_FUSED", "1") != "0"
        # row-sparse token-table gradient: the optimiser applies it to the <= N*S touched rows only (csrc/optim_kernels.cu);
        # the trainer switches it off when gradients are all-reduced densely (data-parallel centralised training)
        self.emb_sparse = os.environ.get("HFL_LM_EMB_SPARSE", "1") != "0"
        self.emb_src: Optional[torch.Tensor] = None      # int32 [N*S]: table row read by every token of the last step
        self.keep_last = False                     # numerics tests: keep the forward intermediates of the last step
        self.last: Dict = {}

    # -- buffers ------------------------------------------------------------------------------------------
    def _grad_buffers(self, P: Dict[str, torch.Tensor]) -> None:
        """fp32 gradient pools: ``small`` = every 1-D parameter (biases, LayerNorm), ``big`` = token + position
        tables (scatter-add targets).  One memset each per step."""
        if self._small is not None:
            return
        dev = P[EMB].device
        small = [(k, v.numel()) for k, v in P.items() if v.dim() == 1]
        big = [(k, P[k].numel()) for k in (EMB, POS)]
        self._small = torch.zeros(sum(n for _, n in small), dtype=torch.float32, device=dev)
        self._big = torch.zeros(sum(n for _, n in big), dtype=torch.float32, device=dev)
        o = 0
        for k, n in small:
            self._small_views[k] = self._small[o:o + n]
            o += n
        o = 0
        for k, n in big:
            self._big_views[k] = self._big[o:o + n].view(P[k].shape)
            o += n

    # -- GEMM helpers (tcgen05 when the operands sit on 16-byte boundaries, torch otherwise) ---------------------
    @staticmethod
    def _lin(x, w, b, alpha, out=None):
        if _aligned(x) and _aligned(w):
            return cu.C().gemm_bf16_tn(x, w, b, float(alpha), False, False, True, out)
        y = torch.addmm(b, x, w.t())
        y = y if alpha == 1.0 else y * alpha
        if out is not None:
            out.copy_(y)
            return out
        return y

    def _lin_ffn(self, x, w, b, alpha):
        """``_lin`` for the feed-forward GEMMs: MXFP8 operands when enabled and K is a multiple of the 128-element
        scale-factor tile, bf16 otherwise."""
        if self.fp8 and x.shape[1] % 128 == 0 and _aligned(x) and _aligned(w) and x.is_contiguous() and w.is_contiguous():
            from ..ops import mxfp8 as mx
            aq, sfa = mx.quantize(x)
            bq, sfb = mx.quantize(w)
            return mx.mx_gemm(aq, sfa, bq, sfb, b.float() * alpha, alpha, False)
        return self._lin(x, w, b, alpha)

    @staticmethod
    def _dx(dy, w, alpha):
        """``alpha * dy @ w`` (w read MN-major in place)."""
        if _aligned(dy) and _aligned(w):
            return cu.C().gemm_bf16_general(dy, False, w, True, float(alpha), False)
        y = dy @ w
        return y if alpha == 1.0 else y * alpha

    @staticmethod
    def _dw(dy, x, alpha):
        """``alpha * dy^T @ x`` (both operands read MN-major in place)."""
        if _aligned(dy) and _aligned(x):
            return cu.C().gemm_bf16_general(dy, True, x, True, float(alpha), False)
        y = dy.t() @ x
        return y if alpha == 1.0 else y * alpha

    # -- evaluation: forward only ------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_hidden_eval(self, P: Dict[str, torch.Tensor], tokens: torch.Tensor, seed_t: torch.Tensor,
                            step_t: torch.Tensor) -> torch.Tensor:
        """Eval-mode forward (Scaler = identity, no dropout, 15 % [MASK] replacement as in every forward of the
        reference, ``src/models/transformer.py:149-152``) up to the decoder LayerNorm on the fused kernels: returns
        ``d`` ``[N*S, E]``; the caller applies the vocabulary projection / loss.  ~30 launches for any batch size instead
        of the ~120 eager ops of the functional model."""
        C = cu.C()
        N, S = tokens.shape
        H, V, L = self.H, self.V, self.L
        E = P[EMB].shape[1]
        att_scale = 1.0 / math.sqrt(E // H)
        tok = tokens.reshape(-1).contiguous()
        x = C.lm_embed_fwd(tok, P[EMB], P[POS], P["transformer_embedding.norm.weight"],
                           P["transformer_embedding.norm.bias"], S, V, LN_EPS, self.mask_rate, 0.0, seed_t, step_t, 1)[0]
        for l in range(L):
            pre = "transformer_encoder.layers.%d" % l
            q = self._lin(x, P[pre + ".mha.linear_q.weight"], P[pre + ".mha.linear_q.bias"], 1.0)
            k = self._lin(x, P[pre + ".mha.linear_k.weight"], P[pre + ".mha.linear_k.bias"], 1.0)
            v = self._lin(x, P[pre + ".mha.linear_v.weight"], P[pre + ".mha.linear_v.bias"], 1.0)
            o = C.lm_attn_fwd(q, k, v, N, S, H, att_scale)[0]
            oo = self._lin(o, P[pre + ".mha.linear_o.weight"], P[pre + ".mha.linear_o.bias"], 1.0)
            x1 = C.lm_add_ln_fwd(x, oo, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"], LN_EPS, 0.0, seed_t, step_t, 0)[0]
            u = self._lin(x1, P[pre + ".linear1.weight"], P[pre + ".linear1.bias"], 1.0)
            hh = C.lm_gelu_drop_fwd(u, 0.0, seed_t, step_t, 0)[0]
            f = self._lin(hh, P[pre + ".linear2.weight"], P[pre + ".linear2.bias"], 1.0)
            x = C.lm_add_ln_fwd(x1, f, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"], LN_EPS, 0.0, seed_t, step_t, 0)[0]
        du = self._lin(x, P["decoder.linear1.weight"], P["decoder.linear1.bias"], 1.0)
        dg = C.lm_gelu_drop_fwd(du, 0.0, seed_t, step_t, 0)[0]
        return C.lm_add_ln_fwd(None, dg, P["decoder.norm1.weight"], P["decoder.norm1.bias"], LN_EPS, 0.0, seed_t, step_t, 0)[0]

    # -- one step --------------------------------------------------------------------------------------------
    def train_step(self, P: Dict[str, torch.Tensor], tokens: torch.Tensor, label_mask: Optional[torch.Tensor],
                   meter: Optional[Tuple[torch.Tensor, float]], seed_t: torch.Tensor, step_t: torch.Tensor,
                   train: bool = True) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        C = cu.C()
        self._grad_buffers(P)
        N, S = tokens.shape
        M = N * S
        H, V, L = self.H, self.V, self.L
        E = P[EMB].shape[1]
        dh = E // H
        s = self.scale_rate if train else 1.0
        inv = 1.0 / s
        p = self.dropout if train else 0.0
        att_scale = inv * inv / math.sqrt(dh)
        tok = tokens.reshape(-1).contiguous()
        G: Dict[str, torch.Tensor] = {}
        self._small.zero_()
        if self.emb_sparse:
            self._big_views[POS].zero_()      # the token-table rows are re-zeroed by the sparse optimiser pass
        else:
            self._big.zero_()
        sg = self._small_views

        # ---- forward ---------------------------------------------------------------------------------------
        x, src, xh0, rstd0, dm0 = C.lm_embed_fwd(tok, P[EMB], P[POS], P["transformer_embedding.norm.weight"],
                                                 P["transformer_embedding.norm.bias"], S, V, LN_EPS * s * s,
                                                 self.mask_rate, p, seed_t, step_t, 1)
        saved: List[Dict] = []
        for l in range(L):
            pre = "transformer_encoder.layers.%d" % l
            site = 16 * (l + 1)
            Wq, Wk, Wv = (P[pre + ".mha.linear_%s.weight" % n] for n in "qkv")
            grouped = self.grouped and _aligned(x) and all(_aligned(w) for w in (Wq, Wk, Wv)) and \
                Wq.stride(0) == Wk.stride(0) == Wv.stride(0)
            if grouped:     # q, k, v share the A operand: ONE launch (csrc/gemm_grouped_tcgen05.cu)
                q, k, v = C.gemm_grouped([x], False, [Wq, Wk, Wv], False,
                                         [P[pre + ".mha.linear_%s.bias" % n] for n in "qkv"], [1.0, 1.0, inv], False)
            else:
                q = self._lin(x, Wq, P[pre + ".mha.linear_q.bias"], 1.0)
                k = self._lin(x, Wk, P[pre + ".mha.linear_k.bias"], 1.0)
                v = self._lin(x, Wv, P[pre + ".mha.linear_v.bias"], inv)
            o, pr = C.lm_attn_fwd(q, k, v, N, S, H, att_scale)
            oo = self._lin(o, P[pre + ".mha.linear_o.weight"], P[pre + ".mha.linear_o.bias"], inv)
            x1, xh1, rstd1, dm1 = C.lm_add_ln_fwd(x, oo, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"], LN_EPS, p,
                                                  seed_t, step_t, site)
            u = self._lin_ffn(x1, P[pre + ".linear1.weight"], P[pre + ".linear1.bias"], inv)
            hh, dmg = C.lm_gelu_drop_fwd(u, p, seed_t, step_t, site + 1)
            f = self._lin_ffn(hh, P[pre + ".linear2.weight"], P[pre + ".linear2.bias"], inv)
            x2, xh2, rstd2, dm2 = C.lm_add_ln_fwd(x1, f, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"], LN_EPS, p,
                                                  seed_t, step_t, site + 2)
            saved.append(dict(x=x, q=q, k=k, v=v, o=o, pr=pr, x1=x1, xh1=xh1, rstd1=rstd1, dm1=dm1, u=u, hh=hh,
                              dmg=dmg, xh2=xh2, rstd2=rstd2, dm2=dm2, grouped=grouped))
            x = x2
        du = self._lin(x, P["decoder.linear1.weight"], P["decoder.linear1.bias"], inv)
        dg, _ = C.lm_gelu_drop_fwd(du, 0.0, seed_t, step_t, 0)
        d, xhd, rstdd, _ = C.lm_add_ln_fwd(None, dg, P["decoder.norm1.weight"], P["decoder.norm1.bias"], LN_EPS, 0.0,
                                           seed_t, step_t, 0)
        n_weight = float(N) if meter is None else float(meter[1])
        W2, b2 = P["decoder.linear2.weight"], P["decoder.linear2.bias"]
        if self.vocab_fused and C.lm_vocab_supported(M, V, E) and _aligned(d) and _aligned(W2) and b2.is_contiguous():
            # ---- fused head: logits live only in TMEM; loss, dW2, db2 and dX come out of three launches ----------
            loss, G["decoder.linear2.weight"], dd32 = C.lm_vocab_head(
                d, W2, b2, tok, label_mask, n_weight, None if meter is None else meter[0], sg["decoder.linear2.bias"], True)
            if self.keep_last:
                self.last = dict(src=src, dm0=dm0, saved=saved, logits=None, x_final=x)
            dzd, _ = C.lm_add_ln_bwd([], xhd, rstdd, P["decoder.norm1.weight"], dm0, 0.0, sg["decoder.norm1.weight"],
                                     sg["decoder.norm1.bias"], None, 0.0, dd32)
        else:
            Vp = (V + 7) // 8 * 8                       # row-padded logits: 16-byte row strides for the TMA maps
            buf = self._logits.get(M)
            if buf is None:
                buf = torch.zeros(M, Vp, dtype=torch.bfloat16, device=tok.device)
                self._logits[M] = buf
            logits = self._lin(d, W2, b2, 1.0, out=buf[:, :V])
            loss, dlogits = C.masked_ce(logits, tok, label_mask, n_weight, None if meter is None else meter[0], True)
            if self.keep_last:
                self.last = dict(src=src, dm0=dm0, saved=saved, logits=logits, x_final=x)

            # ---- backward --------------------------------------------------------------------------------------
            G["decoder.linear2.weight"] = self._dw(dlogits, d, 1.0)
            C.lm_colsum(dlogits, 1.0, sg["decoder.linear2.bias"])
            dd = self._dx(dlogits, W2, 1.0)
            dzd, _ = C.lm_add_ln_bwd([dd], xhd, rstdd, P["decoder.norm1.weight"], dm0, 0.0, sg["decoder.norm1.weight"],
                                     sg["decoder.norm1.bias"], None, 0.0)
        ddu = C.lm_gelu_drop_bwd(dzd, du, dm0, 0.0, sg["decoder.linear1.bias"], inv)
        G["decoder.linear1.weight"] = self._dw(ddu, x, inv)
        parts = [self._dx(ddu, P["decoder.linear1.weight"], inv)]
        for l in reversed(range(L)):
            pre = "transformer_encoder.layers.%d" % l
            sv = saved[l]
            dz2, df = C.lm_add_ln_bwd(parts, sv["xh2"], sv["rstd2"], P[pre + ".norm2.weight"], sv["dm2"], p,
                                      sg[pre + ".norm2.weight"], sg[pre + ".norm2.bias"], sg[pre + ".linear2.bias"], inv)
            G[pre + ".linear2.weight"] = self._dw(df, sv["hh"], inv)
            dhh = self._dx(df, P[pre + ".linear2.weight"], inv)
            dul = C.lm_gelu_drop_bwd(dhh, sv["u"], sv["dmg"], p, sg[pre + ".linear1.bias"], inv)
            G[pre + ".linear1.weight"] = self._dw(dul, sv["x1"], inv)
            dx1 = self._dx(dul, P[pre + ".linear1.weight"], inv)
            dz1, doo = C.lm_add_ln_bwd([dz2, dx1], sv["xh1"], sv["rstd1"], P[pre + ".norm1.weight"], sv["dm1"], p,
                                       sg[pre + ".norm1.weight"], sg[pre + ".norm1.bias"],
                                       sg[pre + ".mha.linear_o.bias"], inv)
            G[pre + ".mha.linear_o.weight"] = self._dw(doo, sv["o"], inv)
            do = self._dx(doo, P[pre + ".mha.linear_o.weight"], inv)
            dq, dk, dv = C.lm_attn_bwd(do, sv["q"], sv["k"], sv["v"], sv["o"], sv["pr"], N, S, H, att_scale,
                                       sg[pre + ".mha.linear_q.bias"], sg[pre + ".mha.linear_k.bias"],
                                       sg[pre + ".mha.linear_v.bias"], inv, bool(sv["grouped"]))
            xin = sv["x"]
            if sv["grouped"]:
                # dv arrives pre-multiplied by 1/s: the three weight gradients are one launch (shared B operand = x,
                # both operands MN-major) and the three input gradients accumulate in ONE TMEM tile
                Wq, Wk, Wv = (P[pre + ".mha.linear_%s.weight" % n] for n in "qkv")
                gq, gk, gv = C.gemm_grouped([dq, dk, dv], True, [xin], True, None, [1.0], False)
                G[pre + ".mha.linear_q.weight"], G[pre + ".mha.linear_k.weight"] = gq, gk
                G[pre + ".mha.linear_v.weight"] = gv
                parts = [dz1, C.gemm_grouped([dq, dk, dv], False, [Wq, Wk, Wv], True, None, [1.0], True)[0]]
            else:
                G[pre + ".mha.linear_q.weight"] = self._dw(dq, xin, 1.0)
                G[pre + ".mha.linear_k.weight"] = self._dw(dk, xin, 1.0)
                G[pre + ".mha.linear_v.weight"] = self._dw(dv, xin, inv)
                parts = [dz1, self._dx(dq, P[pre + ".mha.linear_q.weight"], 1.0),
                         self._dx(dk, P[pre + ".mha.linear_k.weight"], 1.0),
                         self._dx(dv, P[pre + ".mha.linear_v.weight"], inv)]
        C.lm_embed_bwd(parts, dm0, p, xh0, rstd0, P["transformer_embedding.norm.weight"], src, S,
                       self._big_views[EMB], self._big_views[POS], sg["transformer_embedding.norm.weight"],
                       sg["transformer_embedding.norm.bias"])
        for k_, v_ in sg.items():
            G[k_] = v_
        G[EMB], G[POS] = self._big_views[EMB], self._big_views[POS]
        self.emb_src = src
        return loss.view(()), G
