"""Fused masked-LM kernels (csrc/lm_kernels.cu) and the hand-scheduled transformer step vs plain PyTorch fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

from hfl_b200.config.control import make_cfg

pytestmark = pytest.mark.gpu


def _cu():
    from hfl_b200.ops import cuda as cu
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    assert cu.available(), "extension not built"
    return cu


def _rng(dev, seed=5, step=3):
    return (torch.tensor([seed], dtype=torch.int64, device=dev), torch.tensor([step], dtype=torch.int64, device=dev))


def _close(a, b, tol, what=""):
    a, b = a.float(), b.float()
    err = float((a - b).abs().max())
    ref = float(b.abs().max()) + 1e-6
    assert err <= tol * max(ref, 1.0), (what, err, ref)


@pytest.mark.parametrize("E,S,N", [(256, 64, 1), (128, 64, 3), (16, 36, 2)])
def test_embed_ln_dropout_fwd_bwd(E, S, N):
    C = _cu().C()
    dev = torch.device("cuda")
    torch.manual_seed(E + S)
    V, M = 300, N * S
    emb = torch.randn(V + 1, E, device=dev).bfloat16()
    pos = torch.randn(64, E, device=dev).bfloat16()
    gamma = (1 + 0.1 * torch.randn(E, device=dev)).bfloat16()
    beta = (0.1 * torch.randn(E, device=dev)).bfloat16()
    tok = torch.randint(0, V, (M,), device=dev)
    seed, step = _rng(dev)
    eps, p = 1e-5 * 0.25, 0.2
    out, src, xhat, rstd, dmask = C.lm_embed_fwd(tok, emb, pos, gamma, beta, S, V, eps, 0.15, p, seed, step, 1)
    src = src.long()
    replaced = src != tok
    assert bool(((src == tok) | (src == V)).all())
    if M >= 128:
        assert 0.03 < replaced.float().mean().item() < 0.35
    assert 0.6 < dmask.float().mean().item() < 0.95
    # a different step draws different masks; the same (seed, step) reproduces them
    out2, src2, _, _, dm2 = C.lm_embed_fwd(tok, emb, pos, gamma, beta, S, V, eps, 0.15, p, seed, step + 1, 1)
    assert not torch.equal(dm2, dmask)
    out3, src3, _, _, dm3 = C.lm_embed_fwd(tok, emb, pos, gamma, beta, S, V, eps, 0.15, p, seed, step, 1)
    assert torch.equal(dm3, dmask) and torch.equal(src3.long(), src) and torch.equal(out3, out)
    # fp32 reference with the kernel's own draws
    e32, p32 = emb.float().requires_grad_(True), pos.float().requires_grad_(True)
    g32, b32 = gamma.float().requires_grad_(True), beta.float().requires_grad_(True)
    sidx = torch.arange(M, device=dev) % S
    x = e32[src] + p32[sidx]
    y = F.layer_norm(x, (E,), g32, b32, eps) * dmask.float() / (1 - p)
    _close(out, y, 2e-2, "embed fwd")
    dy = torch.randn(M, E, device=dev).bfloat16()
    dy2 = torch.randn(M, E, device=dev).bfloat16()
    y.backward(dy.float() + dy2.float())
    demb, dpos = torch.zeros(V + 1, E, device=dev), torch.zeros(64, E, device=dev)
    dgam, dbet = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
    C.lm_embed_bwd([dy, dy2], dmask, p, xhat, rstd, gamma, src.int(), S, demb, dpos, dgam, dbet)
    _close(demb, e32.grad, 3e-2, "demb")
    _close(dpos, p32.grad, 3e-2, "dpos")
    _close(dgam, g32.grad, 3e-2, "dgamma")
    _close(dbet, b32.grad, 3e-2, "dbeta")


@pytest.mark.parametrize("E,M,p,res", [(256, 64, 0.2, True), (128, 200, 0.0, True), (32, 40, 0.2, False)])
def test_add_ln_fwd_bwd(E, M, p, res):
    C = _cu().C()
    dev = torch.device("cuda")
    torch.manual_seed(E + M)
    x = torch.randn(M, E, device=dev).bfloat16() if res else None
    h = torch.randn(M, E, device=dev).bfloat16()
    gamma = (1 + 0.1 * torch.randn(E, device=dev)).bfloat16()
    beta = (0.1 * torch.randn(E, device=dev)).bfloat16()
    seed, step = _rng(dev)
    y, xhat, rstd, dmask = C.lm_add_ln_fwd(x, h, gamma, beta, 1e-5, p, seed, step, 7)
    h32 = h.float().requires_grad_(True)
    x32 = x.float().requires_grad_(True) if res else None
    g32, b32 = gamma.float().requires_grad_(True), beta.float().requires_grad_(True)
    hd = h32 * dmask.float() / (1 - p) if p > 0 else h32
    z = (x32 + hd) if res else hd
    ref = F.layer_norm(z, (E,), g32, b32, 1e-5)
    _close(y, ref, 3e-2, "add_ln fwd")
    dys = [torch.randn(M, E, device=dev).bfloat16() for _ in range(3)]
    ref.backward(sum(d.float() for d in dys))
    dgam, dbet, dbias = torch.zeros(E, device=dev), torch.zeros(E, device=dev), torch.zeros(E, device=dev)
    dz, dh = C.lm_add_ln_bwd(dys, xhat, rstd, gamma, dmask, p, dgam, dbet, dbias, 0.5)
    if res:
        _close(dz, x32.grad, 3e-2, "dz")
    _close(dh, h32.grad, 3e-2, "dh")
    _close(dgam, g32.grad, 3e-2, "dgamma")
    _close(dbet, b32.grad, 3e-2, "dbeta")
    _close(dbias, 0.5 * h32.grad.sum(0), 3e-2, "dbias")


@pytest.mark.parametrize("M,Fd,p", [(64, 512, 0.2), (100, 72, 0.0)])
def test_gelu_dropout_fwd_bwd(M, Fd, p):
    C = _cu().C()
    dev = torch.device("cuda")
    torch.manual_seed(M)
    u = torch.randn(M, Fd, device=dev).bfloat16()
    seed, step = _rng(dev)
    h, dmask = C.lm_gelu_drop_fwd(u, p, seed, step, 3)
    u32 = u.float().requires_grad_(True)
    ref = F.gelu(u32)
    if p > 0:
        ref = ref * dmask.view(M, Fd).float() / (1 - p)
        assert 0.7 < dmask.float().mean().item() < 0.9
    _close(h, ref, 2e-2, "gelu fwd")
    dh = torch.randn(M, Fd, device=dev).bfloat16()
    ref.backward(dh.float())
    db = torch.zeros(Fd, device=dev)
    du = C.lm_gelu_drop_bwd(dh, u, dmask, p, db, 2.0)
    _close(du, u32.grad, 2e-2, "du")
    _close(db, 2.0 * u32.grad.sum(0), 3e-2, "dbias")
    out = torch.zeros(Fd, device=dev)
    C.lm_colsum(dh, 0.5, out)
    _close(out, 0.5 * dh.float().sum(0), 1e-2, "colsum")


@pytest.mark.parametrize("N,S,H,dh", [(1, 64, 8, 32), (3, 64, 8, 16), (2, 36, 8, 2), (2, 64, 4, 8),
                                      # >= 32 (row, head) pairs: the tensor-core (mma.sync) kernels are dispatched
                                      (6, 64, 8, 32), (5, 64, 8, 16), (16, 36, 8, 2), (9, 64, 4, 8), (40, 50, 8, 32)])
def test_attention_fwd_bwd(N, S, H, dh):
    C = _cu().C()
    dev = torch.device("cuda")
    torch.manual_seed(N * S + dh)
    E, M = H * dh, N * S
    q, k, v = [torch.randn(M, E, device=dev).bfloat16() for _ in range(3)]
    scale = 0.7 / math.sqrt(dh)
    o, P = C.lm_attn_fwd(q, k, v, N, S, H, scale)
    q32, k32, v32 = [t.float().requires_grad_(True) for t in (q, k, v)]

    def heads(t):
        return t.view(N, S, H, dh).transpose(1, 2)
    att = torch.softmax(heads(q32) @ heads(k32).transpose(-2, -1) * scale, dim=-1)
    ref = (att @ heads(v32)).transpose(1, 2).reshape(M, E)
    _close(P, att, 1e-2, "P")
    _close(o, ref, 2e-2, "o")
    do = torch.randn(M, E, device=dev).bfloat16()
    ref.backward(do.float())
    dbq, dbk, dbv = [torch.zeros(E, device=dev) for _ in range(3)]
    dq, dk, dv = C.lm_attn_bwd(do, q, k, v, o, P, N, S, H, scale, dbq, dbk, dbv, 0.5)
    _close(dq, q32.grad, 3e-2, "dq")
    _close(dk, k32.grad, 3e-2, "dk")
    _close(dv, v32.grad, 3e-2, "dv")
    _close(dbq, q32.grad.sum(0), 3e-2, "dbq")
    _close(dbk, k32.grad.sum(0), 3e-2, "dbk")
    _close(dbv, 0.5 * v32.grad.sum(0), 3e-2, "dbv")


def test_masked_ce_padded_rows():
    """Row-padded logits (ld > V) give the same loss / gradient as the dense layout."""
    C = _cu().C()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    M, V = 64, 33278
    Vp = (V + 7) // 8 * 8
    buf = torch.randn(M, Vp, device=dev).bfloat16()
    logits = buf[:, :V]
    tgt = torch.randint(0, V, (M,), device=dev)
    mask = (torch.rand(V + 1, device=dev) < 0.3).to(torch.uint8)
    mask[tgt] = 1
    loss, dl = C.masked_ce(logits, tgt, mask, float(M), None, True)
    loss2, dl2 = C.masked_ce(logits.contiguous(), tgt, mask, float(M), None, True)
    assert dl.stride(0) == Vp
    assert torch.equal(dl, dl2) and float(loss) == float(loss2)


@pytest.mark.parametrize("rate,N", [(1.0, 1), (0.5, 2)])
def test_fused_transformer_step_matches_autograd(rate, N):
    """Whole-step wiring: with dropout and [MASK] replacement off, loss and every gradient of the hand-scheduled
    backward agree with autograd through the fp32 reference graph (same function the CPU path trains with)."""
    cu = _cu()
    from hfl_b200.models import Model
    from hfl_b200.models.fused_transformer import FusedTransformer
    from hfl_b200.models.transformer import transformer_forward
    from hfl_b200.ops.nn import NormCtx
    from hfl_b200.plan.spec import build_spec
    dev = torch.device("cuda")
    V = 1000
    cfg = make_cfg("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_1", num_tokens=V, classes_size=V)
    spec = build_spec(cfg)
    model = Model(spec, cfg, rate, True)
    torch.manual_seed(3)
    arena = model.new_arena(dev)
    for k, v in model.layout.views(arena).items():            # biases / LN params away from their trivial init
        if v.dim() == 1:
            v.add_(0.1 * torch.randn_like(v))
    wc = arena.bfloat16()
    Pv = model.layout.views(wc)
    tokens = torch.randint(0, V, (N, 64), device=dev)
    mask = (torch.rand(V + 1, device=dev) < 0.5).to(torch.uint8)
    mask[tokens.reshape(-1)] = 1
    mask[V] = 0
    ft = FusedTransformer(spec, model.scale_rate, mask_rate=0.0, dropout=0.0)
    seed, step = _rng(dev)
    meter = torch.zeros(4, dtype=torch.float64, device=dev)
    loss, G = ft.train_step(Pv, tokens, mask, (meter, float(N)), seed, step)
    torch.cuda.synchronize()
    P32 = {k: v.float().detach().requires_grad_(True) for k, v in Pv.items()}
    out = transformer_forward(spec, P32, {"label": tokens, "mlm_mask": torch.zeros(N, 64, dtype=torch.bool, device=dev)},
                              NormCtx("train"), model.scale_rate, mask[:V], False, 0.0, 0.0)
    out["loss"].backward()
    assert float(loss) == pytest.approx(float(out["loss"]), rel=2e-2)
    assert float(meter[2]) == float(N)
    assert set(G) == set(P32)
    worst = 1.0
    for k in P32:
        g, r = G[k].float().reshape(-1), P32[k].grad.reshape(-1)
        assert g.shape == r.shape, k
        if float(r.norm()) < 1e-6:           # e.g. the k bias: softmax is invariant to it, the gradient is round-off
            assert float(g.norm()) < 1e-3, k
            continue
        cos = float(F.cosine_similarity(g, r, dim=0))
        ratio = float(g.norm() / r.norm())
        worst = min(worst, cos)
        assert cos > 0.97 and 0.9 < ratio < 1.1, (k, cos, ratio)
    assert worst > 0.97


def test_fused_transformer_mxfp8_ffn():
    """``HFL_LM_FP8=1``: the feed-forward GEMMs run on block-scaled e4m3 operands; loss and gradients stay within the
    format's error budget of the bf16 step."""
    cu = _cu()
    from hfl_b200.models import Model
    from hfl_b200.models.fused_transformer import FusedTransformer
    from hfl_b200.plan.spec import build_spec
    dev = torch.device("cuda")
    V = 1000
    cfg = make_cfg("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_1", num_tokens=V, classes_size=V)
    spec = build_spec(cfg)
    model = Model(spec, cfg, 1.0, True)
    torch.manual_seed(3)
    arena = model.new_arena(dev)
    Pv = model.layout.views(arena.bfloat16())
    tokens = torch.randint(0, V, (4, 64), device=dev)
    outs = []
    for fp8 in (False, True):
        ft = FusedTransformer(spec, model.scale_rate, mask_rate=0.0, dropout=0.0)
        ft.fp8 = fp8
        seed, step = _rng(dev)
        n0 = cu.launch_count()
        loss, G = ft.train_step(Pv, tokens, None, None, seed, step)
        torch.cuda.synchronize()
        outs.append((float(loss), torch.cat([G[k].float().reshape(-1) for k in sorted(G)]), cu.launch_count() - n0))
    assert outs[1][2] > outs[0][2]                                  # the quantise + block-scaled GEMM launches ran
    assert outs[1][0] == pytest.approx(outs[0][0], rel=2e-2)
    cos = float(F.cosine_similarity(outs[0][1], outs[1][1], dim=0))
    print("loss bf16 %.4f fp8 %.4f, gradient cosine %.4f" % (outs[0][0], outs[1][0], cos))
    assert cos > 0.97


def test_fused_lm_slot_trains_in_graph():
    """TrainSlot picks the fused transformer, captures it in a CUDA graph and the loss falls over replays."""
    _cu()
    from hfl_b200.fed.trainer import TrainSlot
    from hfl_b200.models import Model
    from hfl_b200.models.fused_transformer import FusedTransformer
    from hfl_b200.plan.spec import build_spec
    dev = torch.device("cuda")
    V = 600
    cfg = make_cfg("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_1", num_tokens=V, classes_size=V)
    spec = build_spec(cfg)
    model = Model(spec, cfg, 1.0, True)
    slot = TrainSlot(model, cfg, dev, torch.bfloat16, True)
    assert isinstance(slot.fused, FusedTransformer)
    torch.manual_seed(0)
    slot.load_params(model.new_arena(dev))
    mask = torch.ones(V + 1, dtype=torch.uint8)
    mask[V] = 0
    slot.begin_client(0.1, mask, rng_seed=11)
    base = torch.randint(0, 40, (2, 64), device=dev)
    rows = torch.cat([base.repeat(1, 12), base[:, :20]], 1)       # 12 full windows (graph) + one partial (eager)
    losses = []
    for ep in range(3):
        slot.meter.zero_()
        for _ in slot.train_lm(rows, 1, 64):
            pass
        t = slot.meter.buf.tolist()
        losses.append(t[0] / t[2])
    assert all(math.isfinite(v) for v in losses)
    assert losses[-1] < losses[0] * 0.8, losses
    assert slot.graph_kernels[("lm", 2, 64)] < 200


def test_gemm_bias_inside_padded_out_and_ragged_extents():
    """GEMM features the fused LM step leans on: bf16 bias inside alpha, caller-provided row-padded output, reduction
    and output extents that are not multiples of 8 (V = 33 278) on both kernel families."""
    C = _cu().C()
    dev = torch.device("cuda")
    torch.manual_seed(1)
    M, K, V = 64, 256, 33278
    Vp = (V + 7) // 8 * 8
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(V, K, device=dev) / 16).bfloat16()
    b = torch.randn(V, device=dev).bfloat16()
    buf = torch.full((M, Vp), 7.0, device=dev).bfloat16()
    y = C.gemm_bf16_tn(x, w, b, 0.5, False, False, True, buf[:, :V])
    ref = 0.5 * (x.float() @ w.float().t() + b.float())
    assert y.data_ptr() == buf.data_ptr() and y.stride(0) == Vp
    _close(y, ref, 2e-2, "logits")
    assert bool((buf[:, V:] == 7.0).all())                     # pad columns untouched
    dl = torch.zeros(M, Vp, device=dev).bfloat16()
    dl[:, :V] = (torch.randn(M, V, device=dev) / 8).bfloat16()
    dlv = dl[:, :V]
    d = torch.randn(M, K, device=dev).bfloat16()
    dw = C.gemm_bf16_general(dlv, True, d, True, 1.0, False)    # [V, K] = dl^T d, reduction over M, ragged output rows
    _close(dw, dlv.float().t() @ d.float(), 2e-2, "dW")
    dx = C.gemm_bf16_general(dlv, False, w, True, 1.0, False)   # [M, K] = dl w, ragged reduction (33 278)
    _close(dx, dlv.float() @ w.float(), 3e-2, "dX")


@pytest.mark.gpu
@pytest.mark.parametrize("M,E", [(64, 256), (64, 128), (128, 64), (6400, 256)])
def test_grouped_qkv_gemms(M, E):
    """csrc/gemm_grouped_tcgen05.cu: the three forms used for q/k/v (forward with bias, dW, accumulated dX) vs fp32."""
    from hfl_b200.ops import cuda as cu
    C = cu.C()
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(M, E, device=dev).bfloat16()
    W = [(torch.randn(E, E, device=dev) * 0.05).bfloat16() for _ in range(3)]
    b = [(torch.randn(E, device=dev) * 0.1).bfloat16() for _ in range(3)]
    al = [1.0, 1.0, 0.5]
    q, k, v = C.gemm_grouped([x], False, W, False, b, al, False)
    for y, w, bb, a in zip((q, k, v), W, b, al):
        ref = a * (x.float() @ w.float().t() + bb.float())
        assert torch.allclose(y.float(), ref, atol=3e-2, rtol=3e-2), float((y.float() - ref).abs().max())
    dy = [(torch.randn(M, E, device=dev) * 0.1).bfloat16() for _ in range(3)]
    gw = C.gemm_grouped(dy, True, [x], True, None, [1.0], False)
    for g, d in zip(gw, dy):
        ref = d.float().t() @ x.float()
        assert torch.allclose(g.float(), ref, atol=0.05 * float(ref.abs().max()) + 1e-2, rtol=3e-2)
    dx = C.gemm_grouped(dy, False, W, True, None, [1.0], True)[0]
    ref = sum(d.float() @ w.float() for d, w in zip(dy, W))
    assert torch.allclose(dx.float(), ref, atol=0.03 * float(ref.abs().max()) + 1e-2, rtol=3e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("M,E,V,masked", [(64, 256, 33278, True), (128, 128, 1000, True), (40, 64, 515, False),
                                           (64, 256, 4096, False)])
def test_fused_vocab_head_matches_torch(M, E, V, masked):
    """decoder.linear2 + zero-fill label mask + cross-entropy, forward and backward, without materialised logits
    (csrc/lm_vocab_tcgen05.cu) against the fp32 PyTorch composition of the same ops (reference
    src/models/transformer.py:125-129,154-161)."""
    cu = _cu()
    C = cu.C()
    dev = torch.device("cuda")
    torch.manual_seed(M + E + V)
    assert C.lm_vocab_supported(M, V, E)
    d = torch.randn(M, E, device=dev).bfloat16()
    W = (torch.randn(V, E, device=dev) / E ** 0.5).bfloat16()
    b = (0.1 * torch.randn(V, device=dev)).bfloat16()
    tgt = torch.randint(0, V, (M,), device=dev)
    mask = None
    if masked:
        mask = (torch.rand(V, device=dev) < 0.6).to(torch.uint8)
        mask[tgt[: M // 2]] = 1                         # half of the targets are certainly unmasked
    meter = torch.zeros(4, dtype=torch.float64, device=dev)
    dbias = torch.zeros(V, device=dev)
    loss, dW, dX = C.lm_vocab_head(d, W, b, tgt, mask, float(M), meter, dbias, True)
    torch.cuda.synchronize()
    # fp32 reference
    d32, W32, b32 = (t.float().requires_grad_(True) for t in (d, W, b))
    z = d32 @ W32.t() + b32
    if mask is not None:
        z = z.masked_fill(mask.bool().view(1, -1) == 0, 0.0)
    ref = F.cross_entropy(z, tgt)
    ref.backward()
    assert float(loss) == pytest.approx(float(ref), rel=2e-3, abs=2e-3)
    assert float(meter[0]) == pytest.approx(float(ref) * M, rel=2e-3, abs=2e-3)
    assert float(meter[1]) == float((z.argmax(1) == tgt).sum())
    assert float(meter[2]) == float(M)

    def close(a, r, what):
        err = float((a.float() - r).abs().max())
        assert err < 2e-2 * max(float(r.abs().max()), 1e-6) + 1e-6, (what, err, float(r.abs().max()))
    close(dW, W32.grad, "dW")
    close(dX, d32.grad, "dX")
    close(dbias, b32.grad, "dbias")
    # evaluation mode: loss / meter only
    meter2 = torch.zeros(4, dtype=torch.float64, device=dev)
    loss2, _, _ = C.lm_vocab_head(d, W, b, tgt, mask, float(M), meter2, None, False)
    assert float(loss2) == pytest.approx(float(loss), rel=1e-5)


@pytest.mark.gpu
def test_fused_eval_forward_matches_functional_model():
    """Evaluation forward on the fused LM kernels (encoder up to the decoder LayerNorm) == the functional eval-mode model
    with the [MASK] replacement switched off on both sides."""
    cu = _cu()
    from hfl_b200.models import Model
    from hfl_b200.models.fused_transformer import FusedTransformer
    from hfl_b200.models.transformer import transformer_forward
    from hfl_b200.ops.nn import NormCtx
    from hfl_b200.plan.spec import build_spec
    dev = torch.device("cuda")
    V = 1000
    cfg = make_cfg("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_1", num_tokens=V, classes_size=V)
    spec = build_spec(cfg)
    model = Model(spec, cfg, 1.0, True)
    torch.manual_seed(5)
    arena = model.new_arena(dev)
    for k, v in model.layout.views(arena).items():
        if v.dim() == 1:
            v.add_(0.1 * torch.randn_like(v))
    Pv = model.layout.views(arena.bfloat16())
    N = 24
    tokens = torch.randint(0, V, (N, 64), device=dev)
    ft = FusedTransformer(spec, model.scale_rate, mask_rate=0.0, dropout=0.0)
    seed, step = _rng(dev)
    d = ft.forward_hidden_eval(Pv, tokens, seed, step)
    z = torch.nn.functional.linear(d.float(), Pv["decoder.linear2.weight"].float(), Pv["decoder.linear2.bias"].float())
    P32 = {k: v.float() for k, v in Pv.items()}
    ref = transformer_forward(spec, P32, {"label": tokens, "mlm_mask": torch.zeros(N, 64, dtype=torch.bool, device=dev)},
                              NormCtx("eval"), model.scale_rate, None, "raw", 0.0, 0.0)["logits"]
    torch.cuda.synchronize()
    lf = torch.nn.functional.cross_entropy(z, tokens.reshape(-1))
    lr = torch.nn.functional.cross_entropy(ref.float(), tokens.reshape(-1))
    assert float(lf) == pytest.approx(float(lr), rel=2e-2)
    cos = float(torch.nn.functional.cosine_similarity(z.reshape(-1), ref.float().reshape(-1), dim=0))
    assert cos > 0.995, cos


@pytest.mark.gpu
def test_sparse_embedding_optimizer_matches_dense(monkeypatch):
    """Row-sparse token-table update (dense zero-gradient pass + correction of the touched rows) == the dense update:
    two TrainSlots with the same seed, one per mode, agree after graph-replayed steps (same kernels, same RNG; only the
    fp32 rounding order of the token-table update differs)."""
    _cu()
    from hfl_b200.fed.trainer import TrainSlot
    from hfl_b200.models import Model
    from hfl_b200.plan.spec import build_spec
    dev = torch.device("cuda")
    V = 600
    cfg = make_cfg("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_1", num_tokens=V, classes_size=V)
    spec = build_spec(cfg)
    model = Model(spec, cfg, 1.0, True)
    torch.manual_seed(0)
    arena = model.new_arena(dev)
    mask = torch.ones(V + 1, dtype=torch.uint8)
    mask[V] = 0
    base = torch.randint(0, 40, (1, 64), device=dev)
    rows = base.repeat(1, 6)
    outs = []
    for sparse in ("1", "0"):
        monkeypatch.setenv("HFL_LM_EMB_SPARSE", sparse)
        slot = TrainSlot(model, cfg, dev, torch.bfloat16, True)
        assert slot.fused.emb_sparse == (sparse == "1")
        slot.load_params(arena.clone())
        slot.begin_client(0.1, mask, rng_seed=11)
        for _ in slot.train_lm(rows, 1, 64):
            pass
        torch.cuda.synchronize()
        outs.append((slot.w.clone(), slot.mom.clone(), slot.wc.float().clone(), slot.meter.buf.clone()))
        if sparse == "1":
            assert float(slot.fused._big_views["transformer_embedding.embedding.weight"].abs().max()) == 0.0
    (wa, ma, ca, ta), (wb, mb, cb, tb) = outs
    assert torch.allclose(ta, tb, rtol=2e-3, atol=1e-6)                 # same losses along the way
    dw = (wa - arena).float()
    assert float((wa - wb).abs().max()) <= 1e-5 * max(1.0, float(wa.abs().max())) + 2e-2 * float(dw.abs().max())
    assert float((ma - mb).abs().max()) <= 2e-2 * max(1e-3, float(mb.abs().max()))
    assert float((ca - cb).abs().max()) <= 2e-2 * max(1.0, float(cb.abs().max()))
