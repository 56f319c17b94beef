"""Run each hand-written kernel a few times on representative shapes (target for `ncu -k regex:...`)."""
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.config.control import make_cfg
from hfl_b200.ops import cuda as cu
from hfl_b200.ops import fed_ops
from hfl_b200.ops import fused as fo
from hfl_b200.plan.layout import build_tables
from hfl_b200.plan.spec import build_spec

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
C = cu.C()
reps = 5

if which in ("all", "conv"):
    x = torch.randn(2560, 64, 32, 32, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last)
    for _ in range(reps):
        C.conv_fprop_tc(x, w, None, None, 1)
    x2 = torch.randn(2560, 256, 8, 8, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w2 = (torch.randn(256, 256, 3, 3, device=dev) * 0.02).bfloat16().contiguous(memory_format=torch.channels_last)
    for _ in range(reps):
        C.conv_fprop_tc(x2, w2, None, None, 1)
if which in ("all", "gemm"):
    a = torch.randn(8192, 4096, device=dev).bfloat16()
    b = torch.randn(4096, 4096, device=dev).bfloat16()
    for _ in range(reps):
        C.gemm_bf16_tn(a, b, None, 1.0, False, False)
    a2 = torch.randn(64, 256, device=dev).bfloat16()
    b2 = torch.randn(33278, 256, device=dev).bfloat16()
    for _ in range(reps):
        C.gemm_bf16_tn(a2, b2, None, 1.0, False, False)
if which in ("all", "fed"):
    cfg = make_cfg("CIFAR10", "resnet18", "1_800_0.1_non-iid-2_fix_a2-b8_bn_1_1", classes_size=10)
    spec = build_spec(cfg)
    tables = build_tables(spec, 1.0, [1.0, 0.5], channels_last=True)
    g = torch.randn(tables.global_layout.size, device=dev)
    levels = [0] * 16 + [1] * 64
    arenas = [torch.randn(tables.level_layouts[l].size, device=dev) for l in levels]
    masks = [torch.ones(10, dtype=torch.uint8) for _ in levels]
    wc = [torch.zeros(a.numel(), device=dev, dtype=torch.bfloat16) for a in arenas]
    mom = [torch.zeros_like(a) for a in arenas]
    for _ in range(reps):
        fed_ops.distribute(tables, g, arenas, levels, wc, mom)
        fed_ops.aggregate(tables, g, arenas, levels, masks)
if which in ("all", "norm"):
    for (Cc, H) in ((64, 32), (256, 8)):
        x = torch.randn(10, Cc, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        ga = torch.ones(Cc, device=dev).bfloat16()
        be = torch.zeros(Cc, device=dev).bfloat16()
        for _ in range(reps):
            s, y, m, iv, _ = fo.bn_act_fwd(x, None, ga, be, 1e-5)
            fo.bn_act_bwd(y, None, None, x, m, iv, ga, be, torch.empty_like(ga), torch.empty_like(be))
if which in ("all", "lm"):
    import math
    N_, S_, H_, E_ = 1, 64, 8, 256
    q, k, v, do = [torch.randn(N_ * S_, E_, device=dev).bfloat16() for _ in range(4)]
    ga = torch.ones(E_, device=dev).bfloat16()
    seed = torch.zeros(1, dtype=torch.int64, device=dev)
    dbs = [torch.zeros(E_, device=dev) for _ in range(3)]
    for _ in range(reps):
        o, P = C.lm_attn_fwd(q, k, v, N_, S_, H_, 1.0 / math.sqrt(32))
        C.lm_attn_bwd(do, q, k, v, o, P, N_, S_, H_, 1.0 / math.sqrt(32), dbs[0], dbs[1], dbs[2], 1.0)
        y, xh, rs, dm = C.lm_add_ln_fwd(q, k, ga, ga, 1e-5, 0.2, seed, seed, 3)
        C.lm_add_ln_bwd([do, q], xh, rs, ga, dm, 0.2, dbs[0], dbs[1], dbs[2], 1.0)
if which in ("all", "convtrain"):
    # batch-10 training shapes: residual + BatchNorm statistics in the epilogue; the deep-K layer runs as 8-CTA split-K
    # clusters with the DSMEM reduction when HFL_CONV_KSPLIT=8
    for (Cc, H) in ((64, 32), (512, 4)):
        x = torch.randn(10, Cc, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(Cc, Cc, 3, 3, device=dev) * 0.02).bfloat16().contiguous(memory_format=torch.channels_last)
        res = torch.randn_like(x)
        ws = torch.zeros(2 * Cc + 1, device=dev)
        for _ in range(reps):
            ws.zero_()
            C.conv_fprop_tc(x, w, res, ws, 1)
            C.conv_dgrad_tc(x, w, None, 1)
        dy2 = torch.randn(10, 2 * Cc, H // 2, H // 2, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        w2 = (torch.randn(2 * Cc, Cc, 3, 3, device=dev) * 0.02).bfloat16().contiguous(memory_format=torch.channels_last)
        for _ in range(reps):
            C.conv_dgrad_tc(dy2, w2, None, 2)
if which in ("all", "vocab"):
    M_, E_, V_ = 64, 256, 33278
    d = torch.randn(M_, E_, device=dev).bfloat16()
    W = (torch.randn(V_, E_, device=dev) / 16).bfloat16()
    b = torch.zeros(V_, device=dev).bfloat16()
    tgt = torch.randint(0, V_, (M_,), device=dev)
    db = torch.zeros(V_, device=dev)
    for _ in range(reps):
        C.lm_vocab_head(d, W, b, tgt, None, float(M_), None, db, True)
torch.cuda.synchronize()
print("done", which)
if which in ("all", "mega"):
    # one launch of the resident-training megakernel: a level-a client, 3 steps of batch 10 (target for ncu -k mega_kernel)
    from hfl_b200.data.datasets import fetch_dataset, process_dataset
    from hfl_b200.data.pipeline import VisionAugment
    from hfl_b200.fed.mega_trainer import MegaSlot
    from hfl_b200.models import Model
    cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", synthetic=True,
                   synthetic_size={"train": 200, "test": 50})
    ds = fetch_dataset(cfg)
    process_dataset(cfg, ds)
    spec = build_spec(cfg)
    model = Model(spec, cfg, 1.0, True)
    slot = MegaSlot(model, cfg, dev, 10, grid=int(sys.argv[2]) if len(sys.argv) > 2 else 148)
    slot.load_params(torch.randn(model.layout.size, device=dev) * 0.05)
    slot.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
    aug = VisionAugment(ds["train"], True, dev, torch.bfloat16)
    imgs, labs = ds["train"].img[:30].to(dev), ds["train"].target[:30].to(dev)
    for _ in range(2):
        for _ in slot.train_vision(imgs, labs, aug, 1, 10, torch.Generator().manual_seed(1)):
            pass
    torch.cuda.synchronize()
    slot.check()
torch.cuda.synchronize()
