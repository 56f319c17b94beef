"""Manual-backward fused ResNet == autograd functional model (CPU fp32 oracle; GPU bf16 kernels)."""
import pytest
import torch

from hfl_b200.config.control import make_cfg
from hfl_b200.models import make_model
from hfl_b200.models.fused_resnet import FusedResNet
from hfl_b200.ops.nn import NormCtx


def _setup(rate, scale, dev="cpu", dtype=torch.float32, cl=False, mask=1, norm="bn"):
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1-c1_%s_%d_%d" % (norm, scale, mask), classes_size=10)
    model = make_model(cfg, rate, channels_last=cl)
    torch.manual_seed(0)
    arena = model.new_arena(dev, dtype)
    # make BN affine params non-trivial
    for k, v in model.layout.views(arena).items():
        if ".n" in k or k.startswith("n4"):
            v.add_(torch.randn_like(v) * 0.2)
    return cfg, model, arena


@pytest.mark.parametrize("rate,scale", [(0.25, 1), (1.0, 0)])
def test_fused_train_step_matches_autograd_cpu(rate, scale):
    # float64: the hand-written backward must agree with autograd to round-off (fp32 BN-backward over 6-sample
    # batches is too ill-conditioned for a tight comparison)
    cfg, model, arena = _setup(rate, scale, dtype=torch.float64)
    P = {k: v.detach().requires_grad_() for k, v in model.layout.views(arena).items()}
    x = torch.randn(6, 3, 32, 32, dtype=torch.float64)
    y = torch.randint(0, 10, (6,))
    lm = torch.zeros(10, dtype=torch.uint8)
    lm[[1, 3, 4, 7, 8]] = 1
    lm[y] = 1
    out = model.forward(P, {"img": x, "label": y}, NormCtx("train"), lm)
    out["loss"].backward()
    fused = FusedResNet(model.spec, model.scale_rate)
    loss, G = fused.train_step({k: v.detach() for k, v in P.items()}, x, y, lm)
    assert torch.allclose(loss, out["loss"], atol=1e-9, rtol=1e-9)
    assert set(G) == set(P)
    for k in P:
        assert torch.allclose(G[k], P[k].grad, atol=1e-8, rtol=1e-6), k


@pytest.mark.parametrize("norm,rate", [("gn", 0.25), ("in", 0.0625), ("ln", 0.125), ("gn", 0.0625)])
def test_fused_groupnorm_family_matches_autograd_cpu(norm, rate):
    """GroupNorm(4) / InstanceNorm / LayerNorm sub-networks through the same hand-written backward."""
    cfg, model, arena = _setup(rate, 1, dtype=torch.float64, norm=norm)
    assert FusedResNet.eligible(model.spec) and not FusedResNet.eligible(model.spec, train_only=False)
    P = {k: v.detach().requires_grad_() for k, v in model.layout.views(arena).items()}
    x = torch.randn(4, 3, 32, 32, dtype=torch.float64)
    y = torch.randint(0, 10, (4,))
    out = model.forward(P, {"img": x, "label": y}, NormCtx("train"), None)
    out["loss"].backward()
    loss, G = FusedResNet(model.spec, model.scale_rate).train_step({k: v.detach() for k, v in P.items()}, x, y, None)
    assert torch.allclose(loss, out["loss"], atol=1e-9, rtol=1e-9)
    for k in P:
        assert torch.allclose(G[k], P[k].grad, atol=1e-8, rtol=1e-6), k


@pytest.mark.parametrize("norm", ["bn", "gn"])
def test_fused_padded_partial_batch_cpu(norm):
    """A batch padded to the captured size with ``nvalid`` real samples == the step on the real samples alone."""
    cfg, model, arena = _setup(0.25, 1, dtype=torch.float64, norm=norm)
    P = dict(model.layout.views(arena))
    x = torch.randn(6, 3, 32, 32, dtype=torch.float64)
    y = torch.randint(0, 10, (6,))
    fused = FusedResNet(model.spec, model.scale_rate)
    loss_a, Ga = fused.train_step(P, x[:4], y[:4], None)
    xp = x.clone()
    xp[4:] = x[:1]                                   # padding slots repeat a real sample (trainer convention)
    loss_b, Gb = fused.train_step(P, xp, y, None, nvalid=torch.tensor([4], dtype=torch.int32))
    assert torch.allclose(loss_a, loss_b, atol=1e-10)
    for k in Ga:
        assert torch.allclose(Ga[k], Gb[k], atol=1e-9, rtol=1e-7), k


def test_fused_stats_and_eval_match_functional_cpu():
    cfg, model, arena = _setup(0.25, 1)
    P = model.layout.views(arena)
    B, Gn = 5, 3
    x = torch.randn(B * Gn, 3, 32, 32)
    fused = FusedResNet(model.spec, model.scale_rate)
    with torch.no_grad():
        c1, c2 = NormCtx("stats", group=B), NormCtx("stats", group=B)
        model.forward(P, {"img": x}, c1, None, want_score="raw")
        fused.forward_stats(P, x, c2)
        c1.n_groups = c2.n_groups = Gn
        r1, r2 = c1.finalize(), c2.finalize()
        assert set(r1) == set(r2)
        for k in r1:
            assert torch.allclose(r1[k][0], r2[k][0], atol=1e-5, rtol=1e-4), k
            assert torch.allclose(r1[k][1], r2[k][1], atol=1e-5, rtol=1e-4), k
        z1 = model.forward(P, {"img": x}, NormCtx("eval", running=r1), None, want_score="raw")["logits"]
        z2 = fused.forward_eval(P, x, fused.eval_affine(P, r2))
        assert torch.allclose(z1, z2, atol=1e-4, rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [1.0, 0.25, 0.0625])
def test_fused_train_step_gpu_bf16(rate):
    """bf16 kernels vs the fp32 functional model on the same (bf16-rounded) weights and inputs."""
    from hfl_b200.ops import cuda as cu
    assert cu.available()
    dev = torch.device("cuda")
    cfg, model, arena = _setup(rate, 1, dev, torch.float32, cl=True)
    wc = arena.bfloat16()
    P32 = {k: v.float().detach().requires_grad_() for k, v in model.layout.views(wc).items()}
    P16 = dict(model.layout.views(wc))
    x = torch.randn(10, 3, 32, 32, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (10,), device=dev)
    lm = torch.ones(10, dtype=torch.uint8, device=dev)
    lm[[2, 5]] = 0
    lm[y] = 1
    torch.backends.cudnn.allow_tf32 = False
    out = model.forward(P32, {"img": x.float(), "label": y}, NormCtx("train"), lm)
    out["loss"].backward()
    fused = FusedResNet(model.spec, model.scale_rate)
    meter = torch.zeros(4, dtype=torch.float64, device=dev)
    loss, G = fused.train_step(P16, x, y, lm, (meter, 10.0))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(out["loss"])) < 0.05 * max(1.0, abs(float(out["loss"])))
    assert float(meter[2]) == 10.0
    # gradient agreement (bf16 activations/gradients through 18 layers of batch-10 BatchNorm): whole-model cosine
    # similarity against fp32, and against the same-precision autograd path (library BN kernels)
    Pb = {k: v.detach().clone().requires_grad_() for k, v in P16.items()}
    outb = model.forward(Pb, {"img": x, "label": y}, NormCtx("train"), lm)
    outb["loss"].backward()
    cat = lambda d: torch.cat([d[k].float().flatten() for k in sorted(P32)])
    g_fused = cat(G)
    g_fp32 = cat({k: P32[k].grad for k in P32})
    g_bf16 = cat({k: Pb[k].grad for k in Pb})
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a, b, dim=0))
    c32, cbf, cref = cos(g_fused, g_fp32), cos(g_fused, g_bf16), cos(g_bf16, g_fp32)
    print("cos(fused,fp32)=%.4f cos(fused,autograd-bf16)=%.4f cos(autograd-bf16,fp32)=%.4f" % (c32, cbf, cref))
    assert c32 > min(0.97, cref - 0.02), (c32, cref)      # at least as close to fp32 as the library bf16 path
    assert float((g_fused.norm() - g_fp32.norm()).abs()) < 0.2 * float(g_fp32.norm())
    for k in P32:
        a, b = G[k].float().flatten(), P32[k].grad.flatten()
        assert cos(a, b) > 0.6, (k, cos(a, b))


@pytest.mark.gpu
def test_fused_norm_kernels_exact_shapes():
    """bn_act_fwd / bn_act_bwd kernels vs the torch oracle for every (C, rows) that occurs in resnet18 levels a-e."""
    from hfl_b200.ops import cuda as cu
    from hfl_b200.ops import fused as fo
    assert cu.available()
    dev = torch.device("cuda")
    torch.manual_seed(1)
    for C, H in [(4, 32), (8, 16), (16, 8), (32, 4), (64, 32), (128, 16), (256, 8), (512, 4), (16, 32), (64, 8)]:
        for groups in (1, 3):
            N = 10 * groups
            x = (torch.randn(N, C, H, H, device=dev) * 1.5 + 0.3).bfloat16().contiguous(memory_format=torch.channels_last)
            r = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
            ga = (torch.rand(C, device=dev) + 0.5).bfloat16()
            be = (torch.randn(C, device=dev) * 0.3).bfloat16()
            for res in (None, r):
                s, y, m, iv, vu = fo.bn_act_fwd(x, res, ga, be, 1e-5, groups)
                s2, y2, m2, iv2, vu2 = fo.bn_act_fwd(x.float().cpu(), None if res is None else res.float().cpu(),
                                                      ga.float().cpu(), be.float().cpu(), 1e-5, groups)
                assert torch.allclose(s.float().cpu(), s2.bfloat16().float(), atol=1e-2), (C, H, groups)
                assert torch.allclose(m.cpu(), m2, atol=2e-2, rtol=2e-2), (C, H, groups)
                assert torch.allclose(vu.cpu(), vu2, atol=3e-2, rtol=3e-2), (C, H, groups)
                assert torch.allclose(y.float().cpu(), y2, atol=6e-2, rtol=3e-2), (C, H, groups)
            if groups == 1:
                dy = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
                dr = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
                s, y, m, iv, _ = fo.bn_act_fwd(x, None, ga, be, 1e-5, 1)
                for (d2, dres, hw) in [(None, None, 0), (dr, None, 0), (None, dr, 0), (None, None, H * H)]:
                    dg, db = torch.empty_like(ga), torch.empty_like(be)
                    dyy = dy if hw == 0 else torch.randn(N, C, device=dev).bfloat16()
                    ds = fo.bn_act_bwd(dyy, d2, dres, x, m, iv, ga, be, dg, db, hw=hw, dy_scale=0.5 if hw else 1.0)
                    dg2, db2 = torch.empty(C), torch.empty(C)
                    ds2 = fo.bn_act_bwd(dyy.float().cpu(), None if d2 is None else d2.float().cpu(),
                                        None if dres is None else dres.float().cpu(), x.float().cpu(), m.cpu(), iv.cpu(),
                                        ga.float().cpu(), be.float().cpu(), dg2, db2, hw=hw,
                                        dy_scale=0.5 if hw else 1.0)
                    scale = float(ds2.abs().max()) + 1e-6
                    assert float((ds.float().cpu() - ds2).abs().max()) < 0.03 * scale + 2e-2, (C, H, hw)
                    assert torch.allclose(dg.float().cpu(), dg2, atol=0.05 * float(dg2.abs().max()) + 1e-2), (C, H, hw)
                    assert torch.allclose(db.float().cpu(), db2, atol=0.05 * float(db2.abs().max()) + 1e-2), (C, H, hw)


def _cl(t):
    return t.bfloat16().contiguous(memory_format=torch.channels_last)


@pytest.mark.gpu
def test_bn_kernels_partial_batch_nvalid():
    """Device-side ``nvalid``: kernels on a padded batch == kernels on the real samples alone; padding rows are zero.
    Shapes cover the cluster path, the register-resident wide slices and the split (2-launch) path."""
    from hfl_b200.ops import cuda as cu
    from hfl_b200.ops import fused as fo
    assert cu.available()
    dev = torch.device("cuda")
    torch.manual_seed(3)
    N = 10
    for C, H in [(4, 32), (32, 4), (64, 32), (128, 16), (512, 4), (16, 16)]:
        for nv in (1, 3, 9, 10):
            nvt = torch.tensor([nv], dtype=torch.int32, device=dev)
            x = _cl(torch.randn(N, C, H, H, device=dev) * 1.5 + 0.3)
            r = _cl(torch.randn(N, C, H, H, device=dev))
            ga = (torch.rand(C, device=dev) + 0.5).bfloat16()
            be = (torch.randn(C, device=dev) * 0.3).bfloat16()
            ws = torch.zeros(2 * C + 1, dtype=torch.float32, device=dev)
            for res in (None, r):
                s, y, m, iv, _ = fo.bn_act_fwd(x, res, ga, be, 1e-5, 1, ws=ws, nvalid=nvt)
                s1, y1, m1, iv1, _ = fo.bn_act_fwd(x[:nv].contiguous(memory_format=torch.channels_last),
                                                   None if res is None else res[:nv].contiguous(
                                                       memory_format=torch.channels_last), ga, be, 1e-5, 1)
                assert torch.allclose(m, m1, atol=1e-4, rtol=1e-4), (C, H, nv)
                assert torch.allclose(iv, iv1, atol=1e-3, rtol=1e-3), (C, H, nv)
                assert torch.allclose(y[:nv].float(), y1.float(), atol=2e-2, rtol=2e-2), (C, H, nv)
                assert float(y[nv:].float().abs().sum()) == 0.0
                if res is not None:
                    assert torch.equal(s[:nv], s1) and float(s[nv:].float().abs().sum()) == 0.0
            assert float(ws.abs().sum()) == 0.0                         # self-cleaning scratch protocol intact
            s, y, m, iv, _ = fo.bn_act_fwd(x, None, ga, be, 1e-5, 1, nvalid=nvt)
            dy, dr = _cl(torch.randn(N, C, H, H, device=dev)), _cl(torch.randn(N, C, H, H, device=dev))
            for (d2, dres, hw) in [(None, None, 0), (dr, None, 0), (None, dr, 0), (None, None, H * H)]:
                dg, db = torch.empty_like(ga), torch.empty_like(be)
                dg1, db1 = torch.empty_like(ga), torch.empty_like(be)
                dyy = dy if hw == 0 else torch.randn(N, C, device=dev).bfloat16()
                cut = lambda t: None if t is None else t[:nv].contiguous(
                    memory_format=torch.channels_last if t.dim() == 4 else torch.contiguous_format)
                ds = fo.bn_act_bwd(dyy, d2, dres, x, m, iv, ga, be, dg, db, hw=hw, dy_scale=0.5 if hw else 1.0,
                                   nvalid=nvt)
                ds1 = fo.bn_act_bwd(cut(dyy), cut(d2), cut(dres), cut(x), m, iv, ga, be, dg1, db1, hw=hw,
                                    dy_scale=0.5 if hw else 1.0)
                sc = float(ds1.float().abs().max()) + 1e-6
                assert float((ds[:nv].float() - ds1.float()).abs().max()) < 0.02 * sc + 1e-2, (C, H, nv, hw)
                assert float(ds[nv:].float().abs().sum()) == 0.0
                assert torch.allclose(dg.float(), dg1.float(), atol=0.02 * float(dg1.float().abs().max()) + 1e-2)
                assert torch.allclose(db.float(), db1.float(), atol=0.02 * float(db1.float().abs().max()) + 1e-2)


@pytest.mark.gpu
def test_gn_kernels_match_oracle():
    """gn_act_fwd / gn_act_bwd (csrc/gn_kernels.cu) vs the torch fp32 oracle: GroupNorm(4), InstanceNorm, LayerNorm."""
    from hfl_b200.ops import cuda as cu
    from hfl_b200.ops import fused as fo
    assert cu.available()
    dev = torch.device("cuda")
    torch.manual_seed(4)
    N = 10
    for C, H in [(4, 32), (8, 16), (32, 4), (64, 32), (128, 16), (512, 4)]:
        for kind in ("gn", "in", "ln"):
            G = fo.NORM_GROUPS[kind](C)
            x, r = _cl(torch.randn(N, C, H, H, device=dev) * 1.5 + 0.3), _cl(torch.randn(N, C, H, H, device=dev))
            for pdt in (torch.bfloat16, torch.float32):
                ga = (torch.rand(C, device=dev) + 0.5).to(pdt)
                be = (torch.randn(C, device=dev) * 0.3).to(pdt)
                for res in (None, r):
                    s, y, m, rs = fo.gn_act_fwd(x, res, ga, be, 1e-5, G)
                    s2, y2, m2, rs2 = fo.gn_act_fwd(x.float().cpu(), None if res is None else res.float().cpu(),
                                                    ga.float().cpu(), be.float().cpu(), 1e-5, G)
                    assert torch.allclose(s.float().cpu(), s2.bfloat16().float(), atol=1e-2), (C, H, kind)
                    assert torch.allclose(m.cpu(), m2, atol=2e-2, rtol=2e-2), (C, H, kind)
                    assert torch.allclose(rs.cpu(), rs2, atol=2e-2, rtol=2e-2), (C, H, kind)
                    assert torch.allclose(y.float().cpu(), y2, atol=6e-2, rtol=3e-2), (C, H, kind)
                s, y, m, rs = fo.gn_act_fwd(x, None, ga, be, 1e-5, G)
                dy, dr = _cl(torch.randn(N, C, H, H, device=dev)), _cl(torch.randn(N, C, H, H, device=dev))
                for (d2, dres, hw) in [(None, None, 0), (dr, None, 0), (None, dr, 0), (None, None, H * H)]:
                    dg, db = torch.empty_like(ga), torch.empty_like(be)
                    dyy = dy if hw == 0 else torch.randn(N, C, device=dev).bfloat16()
                    ds = fo.gn_act_bwd(dyy, d2, dres, x, m, rs, ga, be, dg, db, G, hw=hw, dy_scale=0.5 if hw else 1.0)
                    dg2, db2 = torch.empty(C), torch.empty(C)
                    c = lambda t: None if t is None else t.float().cpu()
                    ds2 = fo.gn_act_bwd(c(dyy), c(d2), c(dres), c(x), m.cpu(), rs.cpu(), c(ga), c(be), dg2, db2, G,
                                        hw=hw, dy_scale=0.5 if hw else 1.0)
                    sc = float(ds2.abs().max()) + 1e-6
                    assert float((ds.float().cpu() - ds2).abs().max()) < 0.03 * sc + 2e-2, (C, H, kind, hw)
                    assert torch.allclose(dg.float().cpu(), dg2, atol=0.05 * float(dg2.abs().max()) + 1e-2), (C, H, kind)
                    assert torch.allclose(db.float().cpu(), db2, atol=0.05 * float(db2.abs().max()) + 1e-2), (C, H, kind)


@pytest.mark.gpu
@pytest.mark.parametrize("norm,rate", [("gn", 1.0), ("gn", 0.0625), ("in", 0.25), ("ln", 0.5)])
def test_fused_groupnorm_train_step_gpu_bf16(norm, rate):
    dev = torch.device("cuda")
    cfg, model, arena = _setup(rate, 1, dev, torch.float32, cl=True, norm=norm)
    wc = arena.bfloat16()
    P32 = {k: v.float().detach().requires_grad_() for k, v in model.layout.views(wc).items()}
    P16 = dict(model.layout.views(wc))
    x = _cl(torch.randn(10, 3, 32, 32, device=dev))
    y = torch.randint(0, 10, (10,), device=dev)
    torch.backends.cudnn.allow_tf32 = False
    out = model.forward(P32, {"img": x.float(), "label": y}, NormCtx("train"), None)
    out["loss"].backward()
    loss, G = FusedResNet(model.spec, model.scale_rate).train_step(P16, x, y, None)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(out["loss"])) < 0.05 * max(1.0, abs(float(out["loss"])))
    # same acceptance as the BatchNorm test: at least as close to fp32 as the library bf16 autograd path
    Pb = {k: v.detach().clone().requires_grad_() for k, v in P16.items()}
    model.forward(Pb, {"img": x, "label": y}, NormCtx("train"), None)["loss"].backward()
    cat = lambda d: torch.cat([d[k].float().flatten() for k in sorted(P32)])
    g_fused, g_fp32, g_bf16 = cat(G), cat({k: P32[k].grad for k in P32}), cat({k: Pb[k].grad for k in Pb})
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a, b, dim=0))
    c, cref = cos(g_fused, g_fp32), cos(g_bf16, g_fp32)
    print("%s rate %g: cos(fused, fp32) = %.4f  cos(autograd-bf16, fp32) = %.4f" % (norm, rate, c, cref))
    assert c > min(0.97, cref - 0.02), (c, cref)
    assert float((g_fused.norm() - g_fp32.norm()).abs()) < 0.2 * float(g_fp32.norm())


@pytest.mark.gpu
@pytest.mark.parametrize("norm", ["bn", "gn"])
def test_fused_padded_partial_batch_gpu(norm):
    """Whole fused step, padded batch of 10 with 7 real samples vs the 7-sample step (bf16 kernels both sides) and vs
    the fp32 functional model on the 7 samples: padding must not cost accuracy."""
    dev = torch.device("cuda")
    cfg, model, arena = _setup(0.5, 1, dev, torch.float32, cl=True, norm=norm)
    wc = arena.bfloat16()
    P16 = dict(model.layout.views(wc))
    P32 = {k: v.float().detach().requires_grad_() for k, v in model.layout.views(wc).items()}
    nv = 7
    x = _cl(torch.randn(10, 3, 32, 32, device=dev))
    x[nv:] = x[:1]
    y = torch.randint(0, 10, (10,), device=dev)
    torch.backends.cudnn.allow_tf32 = False
    out = model.forward(P32, {"img": x[:nv].float(), "label": y[:nv]}, NormCtx("train"), None)
    out["loss"].backward()
    fused = FusedResNet(model.spec, model.scale_rate, overlap_wgrad=False)
    m1, m2 = (torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(2))
    la, Ga = fused.train_step(P16, x[:nv].contiguous(memory_format=torch.channels_last), y[:nv].contiguous(), None,
                              (m1, float(nv)))
    lb, Gb = fused.train_step(P16, x, y, None, (m2, 10.0), nvalid=torch.tensor([nv], dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    assert abs(float(la) - float(lb)) < 2e-2 * max(1.0, abs(float(la)))
    assert float(m2[2]) == float(nv) and float(m1[2]) == float(nv)      # meter weighs the padded step by nvalid
    assert float(m1[1]) == float(m2[1])
    cat = lambda d: torch.cat([d[k].float().flatten() for k in sorted(P32)])
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a, b, dim=0))
    ga, gb, gr = cat(Ga), cat(Gb), cat({k: P32[k].grad for k in P32})
    ca, cb, cab = cos(ga, gr), cos(gb, gr), cos(ga, gb)
    print("%s: cos(sliced, fp32) %.4f  cos(padded, fp32) %.4f  cos(sliced, padded) %.4f" % (norm, ca, cb, cab))
    assert cb > ca - 0.02 and cab > 0.95, (ca, cb, cab)
    assert abs(float(gb.norm()) - float(ga.norm())) < 0.1 * float(ga.norm())


def _setup_conv(rate, dev="cpu", dtype=torch.float32, cl=False):
    cfg = make_cfg("MNIST", "conv", "1_10_0.5_iid_fix_a1-e1_bn_1_1", classes_size=10)
    model = make_model(cfg, rate, channels_last=cl)
    torch.manual_seed(0)
    arena = model.new_arena(dev, dtype)
    for k, v in model.layout.views(arena).items():
        if v.dim() == 1:
            v.add_(torch.randn_like(v) * 0.2)        # non-trivial conv biases and BN affine parameters
    return cfg, model, arena


@pytest.mark.parametrize("rate,nv", [(0.25, None), (0.0625, None), (0.125, 4)])
def test_fused_convnet_matches_autograd_cpu(rate, nv):
    """MNIST CNN: manual backward (bias-free conv + BN + ReLU + maxpool chain) == autograd in float64."""
    from hfl_b200.models.fused_conv import FusedConvNet
    cfg, model, arena = _setup_conv(rate, dtype=torch.float64)
    P = {k: v.detach().requires_grad_() for k, v in model.layout.views(arena).items()}
    x = torch.randn(6, 1, 28, 28, dtype=torch.float64)
    y = torch.randint(0, 10, (6,))
    lm = torch.zeros(10, dtype=torch.uint8)
    lm[[0, 2, 5]] = 1
    lm[y] = 1
    n = 6 if nv is None else nv
    out = model.forward(P, {"img": x[:n], "label": y[:n]}, NormCtx("train"), lm)
    out["loss"].backward()
    fused = FusedConvNet(model.spec, model.scale_rate)
    nvt = None if nv is None else torch.tensor([nv], dtype=torch.int32)
    loss, G = fused.train_step({k: v.detach() for k, v in P.items()}, x, y, lm, nvalid=nvt)
    assert torch.allclose(loss, out["loss"], atol=1e-9, rtol=1e-9)
    assert set(G) == set(P)
    for k in P:
        assert torch.allclose(G[k], P[k].grad, atol=1e-8, rtol=1e-6), k


@pytest.mark.gpu
def test_maxpool_kernels_match_torch():
    from hfl_b200.ops import cuda as cu
    from hfl_b200.ops import fused as fo
    assert cu.available()
    dev = torch.device("cuda")
    torch.manual_seed(2)
    for N, C, H, W in [(10, 64, 28, 28), (10, 128, 14, 14), (10, 256, 7, 7), (3, 4, 28, 28), (10, 16, 7, 7), (2, 8, 5, 6)]:
        x = _cl(torch.randn(N, C, H, W, device=dev).relu())          # ReLU output: plenty of tied zeros
        xr = x.float().requires_grad_()
        yr = torch.nn.functional.max_pool2d(xr, 2)
        y = fo.maxpool2_fwd(x)
        assert torch.equal(y.float(), yr.detach())
        dy = _cl(torch.randn_like(yr))
        yr.backward(dy.float())
        dx = fo.maxpool2_bwd(dy, x)
        assert torch.equal(dx.float(), xr.grad), (N, C, H, W)


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [1.0, 0.0625])
def test_fused_convnet_gpu_bf16(rate):
    from hfl_b200.models.fused_conv import FusedConvNet
    dev = torch.device("cuda")
    cfg, model, arena = _setup_conv(rate, dev, torch.float32, cl=True)
    wc = arena.bfloat16()
    P32 = {k: v.float().detach().requires_grad_() for k, v in model.layout.views(wc).items()}
    P16 = dict(model.layout.views(wc))
    x = _cl(torch.randn(10, 1, 28, 28, device=dev))
    y = torch.randint(0, 10, (10,), device=dev)
    torch.backends.cudnn.allow_tf32 = False
    out = model.forward(P32, {"img": x.float(), "label": y}, NormCtx("train"), None)
    out["loss"].backward()
    meter = torch.zeros(4, dtype=torch.float64, device=dev)
    loss, G = FusedConvNet(model.spec, model.scale_rate).train_step(P16, x, y, None, (meter, 10.0))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(out["loss"])) < 0.05 * max(1.0, abs(float(out["loss"])))
    names = [k for k in sorted(P32) if not (k.endswith(".bias") and P32[k].grad.abs().max() < 1e-4)]
    cat = lambda d: torch.cat([d[k].float().flatten() for k in names])
    g_fused, g_fp32 = cat(G), cat({k: P32[k].grad for k in P32})
    c = float(torch.nn.functional.cosine_similarity(g_fused, g_fp32, dim=0))
    print("conv rate %g: cos(fused, fp32) = %.4f" % (rate, c))
    assert c > 0.97
    assert float((g_fused.norm() - g_fp32.norm()).abs()) < 0.2 * float(g_fp32.norm())
