"""Manual-backward fused ResNet == autograd functional model (CPU fp32 oracle; GPU bf16 kernels)."""
import pytest
import torch

from hfl_b200.config.control import make_cfg
from hfl_b200.models import make_model
from hfl_b200.models.fused_resnet import FusedResNet
from hfl_b200.ops.nn import NormCtx


def _setup(rate, scale, dev="cpu", dtype=torch.float32, cl=False, mask=1):
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1-c1_bn_%d_%d" % (scale, mask), classes_size=10)
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
