"""sm_100a kernels vs plain PyTorch fp32 references (run on the B200 box: ``pytest -m gpu``)."""
import numpy as np
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from hfl_b200.config.control import make_cfg
from hfl_b200.plan.layout import build_tables
from hfl_b200.plan.spec import build_spec


def _cu():
    from hfl_b200.ops import cuda as cu
    assert cu.available(), "extension must be loaded on a GPU box"
    return cu


@pytest.fixture(params=[1, 8], ids=["nosplit", "ksplit8"])
def ksplit_cap(request):
    """Run a convolution test without and with the cluster split-K path (8 CTAs per output tile, DSMEM reduction); the
    production policy (csrc/conv_tcgen05.cu:conv_ksplit) only enables it from 4 ranks up."""
    from hfl_b200.ops import cuda as cu
    if not (torch.cuda.is_available() and cu.available()):
        pytest.skip("needs the CUDA extension")
    cu.C().conv_set_ksplit(request.param)
    yield request.param
    cu.C().conv_set_ksplit(-1)


def test_extension_loaded():
    cu = _cu()
    assert cu.launch_count() >= 0


@pytest.mark.parametrize("data,model,control,extra,cl", [
    ("CIFAR10", "resnet18", "1_10_0.5_non-iid-2_fix_a1-b1-c1-d1-e1_bn_1_1", dict(classes_size=10), True),
    ("MNIST", "conv", "1_10_0.5_iid_fix_b1-e1_gn_1_1", dict(classes_size=10), True),
    ("WikiText2", "transformer", "1_10_0.5_iid_fix_a1-b1-c1-d1-e1_none_1_0", dict(num_tokens=331), False),
])
def test_distribute_aggregate_match_torch(data, model, control, extra, cl):
    _cu()
    from hfl_b200.fed.federation import combine_torch, distribute_torch
    from hfl_b200.ops import fed_ops
    cfg = make_cfg(data, model, control, **extra)
    spec = build_spec(cfg)
    rates = sorted(set(cfg["model_rate"]), reverse=True)
    tables = build_tables(spec, cfg["global_model_rate"], rates, channels_last=cl)
    dev = torch.device("cuda")
    g = torch.randn(tables.global_layout.size, device=dev)
    client_rates = [rates[i % len(rates)] for i in range(7)]
    levels = [tables.level_id(r) for r in client_rates]
    # distribute
    ref = [torch.zeros(tables.level_layouts[l].size, device=dev) for l in levels]
    distribute_torch(tables, g, ref, client_rates)
    w = [torch.full((tables.level_layouts[l].size,), 7.0, device=dev) for l in levels]
    wc = [torch.zeros(tables.level_layouts[l].size, device=dev, dtype=torch.bfloat16) for l in levels]
    mom = [torch.ones(tables.level_layouts[l].size, device=dev) for l in levels]
    fed_ops.distribute(tables, g, w, levels, wc, mom)
    for a, b, c, m, l in zip(ref, w, wc, mom, levels):
        lay = tables.level_layouts[l]
        for name in lay.entries:
            assert torch.equal(lay.view(a, name), lay.view(b, name)), name
            assert torch.equal(lay.view(a, name).bfloat16(), lay.view(c, name)), name
            assert float(lay.view(m, name).abs().max()) == 0.0
    # aggregate
    rows = max(int(p.out.local(cfg["global_model_rate"])) for p in spec.params if p.label_rows)
    rng = np.random.RandomState(0)
    masks = []
    for _ in levels:
        m = torch.zeros(rows, dtype=torch.uint8)
        m[torch.as_tensor(rng.choice(rows, size=max(2, rows // 3), replace=False))] = 1
        masks.append(m)
    local = [torch.randn_like(a) for a in w]
    g_ref, g_k = g.clone(), g.clone()
    combine_torch(tables, g_ref, local, client_rates, masks)
    fed_ops.aggregate(tables, g_k, local, levels, masks)
    torch.cuda.synchronize()
    assert torch.allclose(g_ref, g_k, atol=1e-5, rtol=1e-5)
    assert not torch.equal(g_k, g)


def test_aggregate_many_clients_and_fused_reduce():
    """aggregate v2: no client-count limit (reference handles any M, src/fed.py:180-297) + the small all-reduce that
    rides in block 0 of the same launch."""
    _cu()
    from hfl_b200.fed.federation import combine_torch
    from hfl_b200.ops import fed_ops
    cfg = make_cfg("CIFAR10", "resnet18", "1_800_0.5_non-iid-2_fix_c2-e8_bn_1_1", classes_size=10)
    spec = build_spec(cfg)
    rates = sorted(set(cfg["model_rate"]), reverse=True)
    tables = build_tables(spec, cfg["global_model_rate"], rates, channels_last=True)
    dev = torch.device("cuda")
    g = torch.randn(tables.global_layout.size, device=dev)
    M = 300
    client_rates = [rates[(i * 7) % len(rates)] for i in range(M)]
    levels = [tables.level_id(r) for r in client_rates]
    rng = np.random.RandomState(1)
    masks = []
    for _ in levels:
        m = torch.zeros(10, dtype=torch.uint8)
        m[torch.as_tensor(rng.choice(10, size=2, replace=False))] = 1
        masks.append(m)
    local = [torch.randn(tables.level_layouts[l].size, device=dev) for l in levels]
    g_ref, g_k = g.clone(), g.clone()
    combine_torch(tables, g_ref, local, client_rates, masks)
    dt = fed_ops.device_tables(tables, dev)
    slabs = [torch.arange(6, dtype=torch.float32, device=dev) * (i + 1) for i in range(3)]
    red_src = torch.tensor([t.data_ptr() for t in slabs], dtype=torch.int64, device=dev)
    red_dst = torch.zeros(6, dtype=torch.float32, device=dev)
    out_ptr = torch.tensor([g_k.data_ptr()], dtype=torch.int64, device=dev)
    fed_ops.aggregate_ptrs(dt, [a.data_ptr() for a in local], levels, masks, out_ptr, 0, dt.n_global_blocks2, red_src,
                           red_dst)
    torch.cuda.synchronize()
    assert torch.allclose(g_ref, g_k, atol=2e-5, rtol=1e-4)
    assert torch.equal(red_dst, torch.arange(6, dtype=torch.float32, device=dev) * 6)


@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
def test_sgd_clip_step(gdtype):
    cu = _cu()
    from hfl_b200.fed.trainer import sgd_clip_step_torch
    dev = torch.device("cuda")
    n = 8 * 12345
    w = torch.randn(n, device=dev)
    mom = torch.randn(n, device=dev) * 0.1
    lr_t = torch.tensor([0.1], device=dev)
    scratch = torch.zeros(4, device=dev)
    for scale in (0.001, 3.0):                # below and above the clipping threshold
        g = (torch.randn(n, device=dev) * scale / n ** 0.5 * 3).to(gdtype)
        w_ref, m_ref, g_ref = w.clone(), mom.clone(), g.clone()
        wc_ref = torch.zeros(n, device=dev, dtype=torch.bfloat16)
        sgd_clip_step_torch(w_ref, wc_ref, g_ref, m_ref, 0.1, 0.9, 5e-4, 1.0)
        wc = torch.zeros(n, device=dev, dtype=torch.bfloat16)
        cu.sgd_clip_step(w, wc, g, mom, lr_t, 0.9, 5e-4, 1.0, scratch)
        torch.cuda.synchronize()
        assert torch.allclose(w, w_ref, atol=1e-6, rtol=1e-5)
        assert torch.allclose(mom, m_ref, atol=1e-6, rtol=1e-5)
        assert torch.equal(wc, w.bfloat16())
        assert float(g.float().abs().max()) == 0.0
        assert float(scratch[0]) == 0.0 and float(scratch[1]) == 0.0


def test_multi_tensor_sgd_matches_flat():
    cu = _cu()
    from hfl_b200.fed.trainer import sgd_clip_step_torch
    dev = torch.device("cuda")
    sizes = [8 * 37, 4096 * 3 + 8, 16, 8 * 1001]
    offs, n = [], 0
    for sz in sizes:
        offs.append(n)
        n += sz
    w = torch.randn(n, device=dev)
    mom = torch.randn(n, device=dev) * 0.1
    grads = [(torch.randn(sz, device=dev) * 0.05).to(torch.bfloat16 if i % 2 == 0 else torch.float32)
             for i, sz in enumerate(sizes)]
    gflat = torch.cat([g.float() for g in grads])
    w_ref, m_ref = w.clone(), mom.clone()
    wc_ref = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    sgd_clip_step_torch(w_ref, wc_ref, gflat.clone(), m_ref, 0.1, 0.9, 5e-4, 1.0)
    chunk = cu.C().ms_chunk()
    blk = [(i, s0) for i, sz in enumerate(sizes) for s0 in range(0, sz, chunk)]
    seg = torch.tensor([(g.data_ptr(), o, g.numel(), 1 if g.dtype == torch.bfloat16 else 0)
                        for g, o in zip(grads, offs)], dtype=torch.int64, device=dev)
    wc = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    scratch = torch.zeros(4, device=dev)
    cu.C().multi_sgd_clip_step(seg, torch.tensor(blk, dtype=torch.int32, device=dev).reshape(-1, 2), w, wc, mom,
                               torch.tensor([0.1], device=dev), 0.9, 5e-4, 1.0, scratch)
    torch.cuda.synchronize()
    assert torch.allclose(w, w_ref, atol=1e-6, rtol=1e-5) and torch.allclose(mom, m_ref, atol=1e-6, rtol=1e-5)
    assert torch.equal(wc, w.bfloat16()) and float(scratch[0]) == 0.0


@pytest.mark.parametrize("rows,V,dtype,masked", [(10, 10, torch.float32, True), (10, 10, torch.bfloat16, False),
                                                  (64, 33278, torch.bfloat16, True), (37, 1000, torch.float32, True)])
def test_masked_ce(rows, V, dtype, masked):
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    z = (torch.randn(rows, V, device=dev) * 2).to(dtype).requires_grad_(True)
    y = torch.randint(0, V, (rows,), device=dev)
    mask = None
    if masked:
        mask = (torch.rand(V, device=dev) < 0.5).to(torch.uint8)
        mask[y] = 1
    meter = torch.zeros(4, dtype=torch.float64, device=dev)
    loss = cu.masked_ce(z, y, mask, 7.0, meter)
    (loss * 2.0).backward()
    zr = z.detach().float().requires_grad_(True)
    zm = zr if mask is None else zr.masked_fill(mask.bool().logical_not().view(1, -1), 0.0)
    lr = F.cross_entropy(zm, y)
    (lr * 2.0).backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert torch.allclose(loss, lr, atol=1e-4, rtol=1e-4)
    assert torch.allclose(z.grad.float(), zr.grad, atol=tol * (1.0 / rows + 1e-3), rtol=tol * 10)
    m = meter.tolist()
    assert m[0] == pytest.approx(float(lr) * 7.0, rel=1e-4) and m[2] == 7.0
    assert m[1] == float((zm.argmax(1) == y).sum()) and m[3] == pytest.approx(float(torch.exp(lr)) * 7.0, rel=1e-4)


def test_augment_matches_torch():
    cu = _cu()
    from hfl_b200.data.datasets import synth_vision
    from hfl_b200.data.pipeline import VisionAugment
    dev = torch.device("cuda")
    ds = synth_vision("CIFAR10", "train", 64)
    img = ds.img.to(dev)
    idx = torch.randperm(64, device=dev)[:10]
    for dtype in (torch.float32, torch.bfloat16):
        aug = VisionAugment(ds, True, dev, dtype)
        par = aug.draw(10)
        got = aug(img, idx, par)
        cpu_aug = VisionAugment(ds, True, "cpu", torch.float32)
        want = cpu_aug(ds.img, idx.cpu(), par.cpu())
        assert got.shape == (10, 3, 32, 32) and got.is_contiguous(memory_format=torch.channels_last)
        assert torch.allclose(got.float().cpu(), want, atol=2e-2 if dtype == torch.bfloat16 else 1e-5)


def test_train_slot_graph_matches_eager():
    """A captured train-step graph replays exactly what the eager step computes (fp32 compute)."""
    _cu()
    from hfl_b200.data.datasets import synth_vision
    from hfl_b200.data.pipeline import VisionAugment
    from hfl_b200.fed.trainer import TrainSlot
    from hfl_b200.models import Model
    dev = torch.device("cuda")
    # exact-arithmetic convs so that eager and replay are comparable (cuDNN TF32 / autotuned algos otherwise
    # differ between the two runs and SGD amplifies the difference)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, True, False
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1-d1_bn_1_1", classes_size=10)
    spec = build_spec(cfg)
    ds = synth_vision("CIFAR10", "train", 40)
    img, lab = ds.img.to(dev), ds.target.to(dev)
    aug = VisionAugment(ds, True, dev, torch.float32)
    init = None
    outs = []
    for use_graph in (False, True):
        model = Model(spec, cfg, 0.125, True)
        slot = TrainSlot(model, cfg, dev, torch.float32, use_graph)
        if init is None:
            init = model.new_arena(dev)
        slot.load_params(init)
        slot.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
        g = torch.Generator().manual_seed(5)
        with torch.cuda.stream(slot.stream):
            for _ in slot.train_vision(img, lab, aug, 1, 10, g):
                pass
        torch.cuda.synchronize()
        outs.append((slot.w.clone(), slot.metrics()))
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = old
    assert torch.allclose(outs[0][0], outs[1][0], atol=3e-3, rtol=1e-2)   # cuDNN may pick different algos under capture
    assert outs[0][1]["Local-Loss"] == pytest.approx(outs[1][1]["Local-Loss"], rel=1e-3)


def test_train_slot_partial_batch_in_graph():
    """bf16 fused path: a shard that is not a multiple of the batch size trains entirely inside the captured graph
    (device-side ``nvalid``) and lands where the eager loop (real 7-sample last batch) lands."""
    _cu()
    from hfl_b200.data.datasets import synth_vision
    from hfl_b200.data.pipeline import VisionAugment
    from hfl_b200.fed.trainer import TrainSlot
    from hfl_b200.models import Model
    dev = torch.device("cuda")
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1-d1_bn_1_1", classes_size=10)
    spec = build_spec(cfg)
    ds = synth_vision("CIFAR10", "train", 27)
    img, lab = ds.img.to(dev), ds.target.to(dev)
    aug = VisionAugment(ds, True, dev, torch.bfloat16)
    init, outs = None, []
    for use_graph in (False, True):
        model = Model(spec, cfg, 0.25, True)
        slot = TrainSlot(model, cfg, dev, torch.bfloat16, use_graph)
        assert slot.fused is not None
        if init is None:
            init = model.new_arena(dev)
        slot.load_params(init)
        slot.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
        g = torch.Generator().manual_seed(5)
        with torch.cuda.stream(slot.stream):
            steps = sum(1 for _ in slot.train_vision(img, lab, aug, 2, 10, g))
        torch.cuda.synchronize()
        assert steps == 6
        outs.append((slot.w.clone(), slot.metrics()))
    a, b = outs[0][0], outs[1][0]
    d0 = (a - init).float()
    # six bf16 SGD steps through 17 small-batch BatchNorm layers on two different kernel schedules (3-D launch shapes,
    # cuDNN algorithms): the trajectories agree in direction and size, not bit for bit
    c = float(torch.nn.functional.cosine_similarity(d0, (b - init).float(), dim=0))
    print("cos(eager update, graph update) = %.4f" % c)
    assert c > 0.9
    assert float((b - init).norm()) == pytest.approx(float(d0.norm()), rel=0.1)
    assert outs[0][1]["Local-Loss"] == pytest.approx(outs[1][1]["Local-Loss"], rel=5e-2)
    assert outs[0][1]["Local-Accuracy"] == pytest.approx(outs[1][1]["Local-Accuracy"], abs=10.0)


@pytest.mark.parametrize("data,model,control,sizes", [
    ("CIFAR10", "resnet18", "1_10_0.5_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", {"train": 1000, "test": 200}),
    ("MNIST", "conv", "1_10_0.5_iid_fix_a1-e1_bn_1_1", {"train": 1000, "test": 200}),
])
def test_engine_round_gpu(tmp_path, data, model, control, sizes):
    """Convergence smoke on the GPU path (bf16 compute, CUDA graphs, fused kernels): loss falls, metrics finite."""
    _cu()
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.parallel.dist import DistInfo
    dev = torch.device("cuda", 0)
    cfg = make_cfg(data, model, control, synthetic=True, synthetic_size=sizes)
    cfg["num_epochs"] = {"global": 3, "local": 2}
    eng = Engine(cfg, DistInfo(0, 1, 0, dev, None), options=EngineOptions(out_dir=str(tmp_path), quiet=True))
    eng.fit()
    h = eng.logger.history
    assert all(np.isfinite(v) for v in h["train/Local-Loss"] + h["test/Global-Loss"] + h["test/Local-Loss"])
    assert h["train/Local-Loss"][-1] < h["train/Local-Loss"][0]
    assert h["test/Local-Accuracy"][-1] > 100.0 / 10 * 1.5


def test_engine_lm_round_gpu(tmp_path):
    _cu()
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.parallel.dist import DistInfo
    dev = torch.device("cuda", 0)
    cfg = make_cfg("WikiText2", "transformer", "1_10_0.5_iid_fix_a5-b5_none_1_1", synthetic=True,
                   synthetic_size={"train": 100 * 64 * 6, "test": 10 * 64 * 4}, synthetic_vocab=500)
    cfg["num_epochs"] = {"global": 2, "local": 1}
    eng = Engine(cfg, DistInfo(0, 1, 0, dev, None), options=EngineOptions(out_dir=str(tmp_path), quiet=True))
    eng.fit()
    h = eng.logger.history
    assert all(np.isfinite(v) for v in h["train/Local-Loss"] + h["test/Global-Perplexity"])
    assert h["train/Local-Loss"][-1] < h["train/Local-Loss"][0]


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (64, 256, 256), (640, 64, 576), (10240, 64, 27 * 8), (200, 1000, 128),
                                   (64, 33278, 256), (1, 16, 8), (129, 40, 72)])
def test_tcgen05_gemm_matches_torch(M, N, K):
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(M + N + K)
    x = (torch.randn(M, K, device=dev)).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev)
    ref = x.float() @ w.float().t()
    y = cu.C().gemm_bf16_tn(x, w, None, 1.0, False, True)
    torch.cuda.synchronize()
    assert torch.allclose(y, ref, atol=2e-3, rtol=2e-3), float((y - ref).abs().max())
    y2 = cu.C().gemm_bf16_tn(x, w, b, 0.5, True, False)
    ref2 = torch.nn.functional.gelu(ref * 0.5 + b)
    assert torch.allclose(y2.float(), ref2, atol=3e-2, rtol=3e-2)
    # strided (sliced) operands: K-contiguous views into wider buffers
    if K >= 16:
        xb = torch.randn(M, K + 8, device=dev).bfloat16()
        wb = torch.randn(N, K + 16, device=dev).bfloat16()
        y3 = cu.C().gemm_bf16_tn(xb[:, :K], wb[:, 8:8 + K], None, 1.0, False, True)
        assert torch.allclose(y3, xb[:, :K].float() @ wb[:, 8:8 + K].float().t(), atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("N,C,H,K,R,stride", [(10, 64, 32, 64, 3, 1), (10, 64, 32, 128, 3, 2), (10, 128, 16, 128, 3, 1),
                                               (10, 64, 32, 128, 1, 2), (10, 256, 8, 256, 3, 1), (10, 256, 8, 512, 3, 2),
                                               (10, 512, 4, 512, 3, 1), (3, 128, 16, 256, 1, 2), (64, 64, 32, 64, 3, 1),
                                               # sub-64 channel counts (levels b..e): partial 64-blocks are TMA zero fill
                                               (10, 32, 32, 32, 3, 1), (10, 16, 32, 16, 3, 1), (10, 8, 32, 8, 3, 1),
                                               (10, 32, 32, 64, 3, 2), (10, 16, 16, 32, 1, 2), (10, 8, 32, 16, 3, 2),
                                               (7, 128, 8, 64, 3, 1), (10, 96, 16, 160, 3, 1)])
def test_tcgen05_conv_fprop_matches_torch(N, C, H, K, R, stride, ksplit_cap):
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(N + C + K + R)
    x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, C, R, R, device=dev) / (C * R * R) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    if R == 1:
        w = w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    assert cu.C().conv_tc_supported(N, H, H, C, K, R, R, stride)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), None, stride, R // 2)
    y = cu.C().conv_fprop_tc(x, w, None, None, stride)
    torch.cuda.synchronize()
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    err = float((y.float() - ref).abs().max())
    assert err < 3e-2 * max(1.0, float(ref.abs().max())), err
    # fused epilogue: residual add + BatchNorm batch statistics of the output
    res = torch.randn_like(ref).bfloat16().contiguous(memory_format=torch.channels_last)
    stats = torch.zeros(2, K, device=dev)
    y2 = cu.C().conv_fprop_tc(x, w, res, stats, stride)
    ref2 = (ref + res.float())
    assert float((y2.float() - ref2).abs().max()) < 4e-2 * max(1.0, float(ref2.abs().max()))
    yf = y2.float()
    assert torch.allclose(stats[0], yf.sum(dim=(0, 2, 3)), rtol=2e-3, atol=0.5)
    assert torch.allclose(stats[1], (yf * yf).sum(dim=(0, 2, 3)), rtol=2e-3, atol=0.5)


@pytest.mark.parametrize("N,C,H,K,stride,with_res,nvalid", [(10, 64, 32, 64, 1, True, None), (10, 64, 32, 128, 2, False, None),
                                                            (10, 256, 8, 256, 1, True, None), (10, 512, 4, 512, 1, True, 3),
                                                            (10, 16, 32, 16, 1, True, 7), (10, 8, 32, 16, 2, False, None),
                                                            (10, 128, 16, 128, 1, True, 10)])
def test_conv_bn_act_fused_matches_unfused(N, C, H, K, stride, with_res, nvalid, monkeypatch, ksplit_cap):
    """conv -> (residual + batch statistics in the tcgen05 epilogue) -> bn_apply_sums  ==  conv -> bn_act_fwd, also
    with a padded partial batch, over two consecutive calls (the scratch must come back zeroed)."""
    cu = _cu()
    from hfl_b200.ops import fused as fo
    dev = torch.device("cuda")
    torch.manual_seed(N + C + K)
    monkeypatch.setattr(fo, "CONV_TC_MODE", "on")
    x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    OH = H // stride
    res = torch.randn(N, K, OH, OH, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) if with_res else None
    nv = None
    if nvalid is not None:
        nv = torch.tensor([nvalid], dtype=torch.int32, device=dev)     # padding rows keep arbitrary (non-zero) data
    w = (torch.randn(K, C, 3, 3, device=dev) / (C * 9) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    gamma = (1 + 0.1 * torch.randn(K, device=dev)).bfloat16()
    beta = (0.1 * torch.randn(K, device=dev)).bfloat16()
    ws = torch.zeros(2 * K + 1, device=dev)
    for _ in range(2):
        s, y, m, i = fo.conv_bn_act_fwd(x, w, stride, 1, res, gamma, beta, 1e-5, ws, nv)
        torch.cuda.synchronize()
        assert float(ws.abs().max()) == 0.0                       # self-cleaned for the next step
    # fp32 reference of the same op
    nvi = N if nvalid is None else nvalid
    c = torch.nn.functional.conv2d(x.float(), w.float(), None, stride, 1)
    sr = c if res is None else c + res.float()
    sr[nvi:] = 0
    mean = sr[:nvi].mean(dim=(0, 2, 3))
    var = sr[:nvi].var(dim=(0, 2, 3), unbiased=False)
    yr = torch.relu((sr - mean.view(1, -1, 1, 1)) * torch.rsqrt(var + 1e-5).view(1, -1, 1, 1) * gamma.float().view(1, -1, 1, 1)
                    + beta.float().view(1, -1, 1, 1))
    yr[nvi:] = 0
    assert float((s.float() - sr).abs().max()) < 4e-2 * max(1.0, float(sr.abs().max()))
    assert torch.allclose(m.view(-1), mean, atol=2e-2, rtol=2e-2)
    assert torch.allclose(i.view(-1), torch.rsqrt(var + 1e-5), atol=2e-2, rtol=2e-2)
    assert float((y.float() - yr).abs().max()) < 6e-2 * max(1.0, float(yr.abs().max()))
    assert float(y[nvi:].abs().max()) == 0.0 if nvi < N else True


@pytest.mark.parametrize("N,C2,H2,K,s2,nvalid", [(10, 64, 32, 128, 2, None), (10, 128, 16, 256, 2, 6), (10, 16, 32, 32, 2, None),
                                                (10, 256, 8, 512, 2, None), (10, 32, 16, 64, 1, None)])
def test_conv_shortcut_fold_matches_unfused(N, C2, H2, K, s2, nvalid, monkeypatch, ksplit_cap):
    """Transition block: conv2(o2) + shortcut_1x1(o1) + BatchNorm statistics in ONE tcgen05 launch (the shortcut's k-blocks
    are accumulated into the same TMEM tile) == the two convolutions computed separately in fp32."""
    cu = _cu()
    from hfl_b200.ops import fused as fo
    dev = torch.device("cuda")
    torch.manual_seed(N + C2 + K)
    monkeypatch.setattr(fo, "CONV_TC_MODE", "on")
    OH = H2 // s2
    o2 = torch.randn(N, K, OH, OH, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    o1 = torch.randn(N, C2, H2, H2, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, K, 3, 3, device=dev) / (K * 9) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    wsc = (torch.randn(K, C2, 1, 1, device=dev) / C2 ** 0.5).bfloat16()
    wsc = wsc.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    gamma = (1 + 0.1 * torch.randn(K, device=dev)).bfloat16()
    beta = (0.1 * torch.randn(K, device=dev)).bfloat16()
    ws = torch.zeros(2 * K + 1, device=dev)
    nv = None if nvalid is None else torch.tensor([nvalid], dtype=torch.int32, device=dev)
    s, y, m, i = fo.conv_bn_act_fwd(o2, w, 1, 1, None, gamma, beta, 1e-5, ws, nv, shortcut=(o1, wsc, s2))
    torch.cuda.synchronize()
    assert float(ws.abs().max()) == 0.0
    nvi = N if nvalid is None else nvalid
    sr = torch.nn.functional.conv2d(o2.float(), w.float(), None, 1, 1) + torch.nn.functional.conv2d(o1.float(), wsc.float(), None, s2, 0)
    sr[nvi:] = 0
    mean = sr[:nvi].mean(dim=(0, 2, 3))
    var = sr[:nvi].var(dim=(0, 2, 3), unbiased=False)
    yr = torch.relu((sr - mean.view(1, -1, 1, 1)) * torch.rsqrt(var + 1e-5).view(1, -1, 1, 1) * gamma.float().view(1, -1, 1, 1)
                    + beta.float().view(1, -1, 1, 1))
    yr[nvi:] = 0
    assert float((s.float() - sr).abs().max()) < 4e-2 * max(1.0, float(sr.abs().max()))
    assert torch.allclose(m.view(-1), mean, atol=2e-2, rtol=2e-2)
    assert float((y.float() - yr).abs().max()) < 6e-2 * max(1.0, float(yr.abs().max()))
    # and the unfolded path (HFL_CONV_SC_FOLD=0) gives the same block output
    monkeypatch.setattr(fo, "CONV_SC_FOLD", False)
    s2_, y2, _, _ = fo.conv_bn_act_fwd(o2, w, 1, 1, None, gamma, beta, 1e-5, ws, nv, shortcut=(o1, wsc, s2))
    assert float((s2_.float() - s.float()).abs().max()) < 4e-2 * max(1.0, float(sr.abs().max()))


def test_tc_linear_forward_backward():
    _cu()
    from hfl_b200.ops.linear import tc_linear
    dev = torch.device("cuda")
    torch.manual_seed(3)
    x = torch.randn(2, 64, 256, device=dev).bfloat16().requires_grad_(True)
    w = (torch.randn(512, 256, device=dev) * 0.05).bfloat16().requires_grad_(True)
    b = (torch.randn(512, device=dev) * 0.1).bfloat16().requires_grad_(True)
    y = tc_linear(x, w, b, 0.5)
    (y.float() ** 2).sum().backward()
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.linear(xr, wr, br) * 0.5
    (yr ** 2).sum().backward()
    assert torch.allclose(y.float(), yr, atol=3e-2, rtol=3e-2)
    for a, r in ((x, xr), (w, wr), (b, br)):
        assert torch.allclose(a.grad.float(), r.grad, atol=0.05 * float(r.grad.abs().max()), rtol=5e-2)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 192), (64, 512, 640), (200, 136, 72)])
def test_tcgen05_gemm_general_layouts(a_mn, b_mn, M, N, K):
    """MN-major (transposed-in-memory) operands through UMMA descriptors: what dgrad / wgrad GEMMs need."""
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(M * 3 + N * 5 + K)
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    ref = A.float() @ B.float().t()
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    y = cu.C().gemm_bf16_general(a, a_mn, b, b_mn, 1.0, True)
    torch.cuda.synchronize()
    assert torch.allclose(y, ref, atol=3e-3, rtol=3e-3), float((y - ref).abs().max())


@pytest.mark.parametrize("N,C,H,K,R", [(10, 64, 32, 64, 3), (10, 128, 16, 128, 3), (10, 64, 16, 128, 3), (10, 256, 8, 256, 3),
                                        (10, 512, 4, 512, 3), (4, 128, 16, 64, 1), (10, 32, 32, 32, 3), (10, 16, 32, 16, 3),
                                        (10, 8, 16, 8, 3), (10, 32, 16, 64, 3), (10, 64, 16, 32, 3), (6, 16, 8, 24, 1)])
def test_tcgen05_conv_dgrad_matches_autograd(N, C, H, K, R, ksplit_cap):
    """Stride-1 data gradient on the implicit-GEMM kernel (mirrored taps, MN-major filter operand)."""
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(N + C + K)
    x = torch.randn(N, C, H, H, device=dev, requires_grad=True)
    w = (torch.randn(K, C, R, R, device=dev) / (C * R * R) ** 0.5).bfloat16()
    dy = torch.randn(N, K, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    y = torch.nn.functional.conv2d(x, w.float(), None, 1, R // 2)
    y.backward(dy.float())
    wk = w.contiguous(memory_format=torch.channels_last) if R > 1 else w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    dx = cu.C().conv_dgrad_tc(dy, wk, None)
    torch.cuda.synchronize()
    ref = x.grad
    assert float((dx.float() - ref).abs().max()) < 3e-2 * max(1.0, float(ref.abs().max()))
    res = torch.randn_like(ref).bfloat16().contiguous(memory_format=torch.channels_last)
    dx2 = cu.C().conv_dgrad_tc(dy, wk, res)
    assert float((dx2.float() - ref - res.float()).abs().max()) < 4e-2 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("N,C,H,K,R", [(10, 64, 32, 128, 3), (10, 128, 16, 256, 3), (10, 256, 8, 512, 3), (10, 64, 32, 128, 1),
                                        (10, 16, 32, 32, 3), (10, 8, 16, 16, 1), (3, 32, 16, 64, 3)])
def test_tcgen05_conv_dgrad_stride2_matches_autograd(N, C, H, K, R, ksplit_cap):
    """Data gradient of a stride-2 convolution: the four output-parity classes (1 / 2 / 2 / 4 taps for a 3x3 filter, a
    single class for 1x1 with the other three written as zeros) run inside one launch of the implicit-GEMM kernel."""
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(N + C + K + R)
    x = torch.randn(N, C, H, H, device=dev, requires_grad=True)
    w = (torch.randn(K, C, R, R, device=dev) / (C * R * R) ** 0.5).bfloat16()
    y = torch.nn.functional.conv2d(x, w.float(), None, 2, R // 2)
    dy = torch.randn_like(y).bfloat16().contiguous(memory_format=torch.channels_last)
    y.backward(dy.float())
    wk = w.contiguous(memory_format=torch.channels_last) if R > 1 else w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    assert cu.C().conv_dgrad_tc_supported(N, H, H, C, K, R, R, 2)
    dx = cu.C().conv_dgrad_tc(dy, wk, None, 2)
    torch.cuda.synchronize()
    ref = x.grad
    assert dx.shape == ref.shape
    assert float((dx.float() - ref).abs().max()) < 3e-2 * max(1.0, float(ref.abs().max()))
    res = torch.randn_like(ref).bfloat16().contiguous(memory_format=torch.channels_last)
    dx2 = cu.C().conv_dgrad_tc(dy, wk, res, 2)
    assert float((dx2.float() - ref - res.float()).abs().max()) < 4e-2 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("N,C,H,K,R,stride", [(10, 64, 32, 64, 3, 1), (10, 64, 32, 128, 3, 2), (10, 128, 16, 128, 3, 1),
                                               (10, 64, 32, 128, 1, 2), (10, 256, 8, 256, 3, 1), (12, 512, 4, 512, 3, 1),
                                               (8, 256, 8, 512, 3, 2), (10, 32, 32, 32, 3, 1), (10, 16, 32, 16, 3, 1),
                                               (10, 8, 32, 8, 3, 1), (10, 32, 32, 64, 3, 2), (10, 16, 16, 32, 1, 2),
                                               (10, 64, 16, 32, 3, 1), (10, 8, 32, 16, 3, 2),
                                               # partial last image group (4x4 / 2x2 maps at batch 10: 4 / 16 images per brick)
                                               (10, 512, 4, 512, 3, 1), (10, 256, 8, 512, 3, 2), (10, 128, 4, 128, 1, 2), (7, 64, 4, 64, 3, 1)])
def test_tcgen05_conv_wgrad_matches_autograd(N, C, H, K, R, stride):
    """Weight gradient with both operands MN-major (pixels = reduction), split-K with fp32 atomics."""
    cu = _cu()
    dev = torch.device("cuda")
    torch.manual_seed(N + C + K + R)
    x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, C, R, R, device=dev) * 0.05).requires_grad_(True)
    y = torch.nn.functional.conv2d(x.float(), w, None, stride, R // 2)
    dy = torch.randn_like(y).bfloat16().contiguous(memory_format=torch.channels_last)
    y.backward(dy.float())
    if not cu.C().conv_wgrad_tc_supported(N, H, H, C, K, R, R, stride):
        pytest.skip("shape not supported by the wgrad kernel")
    dw = cu.C().conv_wgrad_tc(dy, x, R, stride)
    torch.cuda.synchronize()
    ref = w.grad
    assert dw.shape == ref.shape
    assert float((dw - ref).abs().max()) < 2e-2 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("M,N,K", [(1024, 1024, 256), (1100, 1032, 320), (1152, 33278, 256), (2048, 1024, 1024)])
def test_cluster_multicast_gemm(M, N, K):
    """2-CTA cluster + TMA multicast + 128x256 tiles (csrc/gemm_mc_tcgen05.cu): odd tile counts (padding CTA of the pair),
    ragged N, row-padded output, every epilogue variant; agrees with fp32 torch and with the single-CTA kernel."""
    cu = _cu()
    C = cu.C()
    dev = torch.device("cuda")
    torch.manual_seed(M + K)
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    ref = x.float() @ w.float().t()
    y = C.gemm_bf16_tn(x, w, None, 1.0, False, False, False, None)
    assert torch.allclose(y.float(), ref, atol=3e-2, rtol=3e-2), float((y.float() - ref).abs().max())
    Np = (N + 7) // 8 * 8
    buf = torch.full((M, Np), 3.0, device=dev).bfloat16()
    y2 = C.gemm_bf16_tn(x, w, b, 0.5, True, False, True, buf[:, :N])
    ref2 = torch.nn.functional.gelu(0.5 * (ref + b.float()))
    assert torch.allclose(y2.float(), ref2, atol=3e-2, rtol=3e-2)
    if Np > N:
        assert bool((buf[:, N:] == 3.0).all())
    y3 = C.gemm_bf16_tn(x, w, b.float(), 2.0, False, False, False, None)
    assert torch.allclose(y3.float(), 2.0 * ref + b.float(), atol=6e-2, rtol=3e-2)


