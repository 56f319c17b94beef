"""Multi-GPU paths (needs >= 2 GPUs on one box: ``gpurun --gpus 2 -- pytest -m gpu tests/test_multi_gpu.py``)."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

from hfl_b200.config.control import make_cfg

CONTROL = "1_10_0.6_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1"


def _cfg(control=CONTROL):
    cfg = make_cfg("CIFAR10", "resnet18", control, synthetic=True, synthetic_size={"train": 1000, "test": 200})
    cfg["num_epochs"] = {"global": 2, "local": 1}
    return cfg


def _agg_setup():
    """Deterministic aggregation problem shared by the workers and the oracle."""
    from hfl_b200.plan.layout import build_tables
    from hfl_b200.plan.spec import build_spec
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_non-iid-2_fix_a1-b1-c1-d1-e1_bn_1_1", classes_size=10)
    spec = build_spec(cfg)
    rates = sorted(set(cfg["model_rate"]), reverse=True)
    tables = build_tables(spec, 1.0, rates, channels_last=True)
    levels = [0, 4, 2, 1, 3, 4, 0]
    masks = []
    for i in range(len(levels)):
        m = torch.zeros(10, dtype=torch.uint8)
        m[torch.tensor([(i * 3) % 10, (i * 3 + 1) % 10, (i + 5) % 10])] = 1
        masks.append(m)
    assignment = [[0, 2, 5], [1, 3, 4, 6]]
    return tables, levels, masks, assignment


def _client_arena(tables, level, i):
    g = torch.Generator().manual_seed(100 + i)
    return torch.randn(tables.level_layouts[level].size, generator=g)


def _active(rnd, n):
    return None if rnd % 3 != 1 else [i for i in range(n) if i != rnd % n]


def _agg_worker(rank, world, port, ret, rounds=3, jitter=False):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hfl_b200.parallel.dist import init_distributed, shutdown
    from hfl_b200.parallel.peer import PeerPool, p2p_aggregate
    info = init_distributed()
    tables, levels, masks, assignment = _agg_setup()
    gsize = tables.global_layout.size
    pool = PeerPool(info, (gsize + sum(tables.level_layouts[l].size for l in levels)) * 4 + (1 << 20))
    ga = pool.alloc_persistent(gsize, torch.float32)
    ga.copy_(torch.randn(gsize, generator=torch.Generator().manual_seed(7)).to(info.device))
    pool.exchange()
    gen = torch.Generator().manual_seed(1234 + rank)
    for rnd in range(rounds):                # several rounds: buffers are re-used, barriers re-armed
        pool.reset_round()
        if jitter:                           # ranks drift apart by up to ~2 ms of device time, no host sync between
            torch.cuda._sleep(int(torch.randint(0, 4_000_000, (1,), generator=gen)))   # rounds: only the flag
        for i in assignment[rank]:                                                     # protocol orders them
            a = pool.alloc_round(tables.level_layouts[levels[i]].size, torch.float32)
            a.copy_((_client_arena(tables, levels[i], i) + rnd).to(info.device), non_blocking=True)
        p2p_aggregate(pool, tables, ga, levels, assignment, masks, info,
                      active=_active(rnd, len(levels)) if jitter else None)
        if not jitter:
            torch.cuda.synchronize()
            pool.check()
    torch.cuda.synchronize()
    pool.check()
    torch.save(ga.float().cpu().clone(), ret + ".%d" % rank)
    torch.distributed.barrier()
    pool.close()
    shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_p2p_aggregate_exact(tmp_path):
    """The fused P2P kernel (peer loads + peer stores + flag barriers) == the torch oracle on every rank's replica."""
    from hfl_b200.fed.federation import combine_torch
    tables, levels, masks, assignment = _agg_setup()
    ret = str(tmp_path / "ga.pt")
    mp.spawn(_agg_worker, args=(2, 29650 + os.getpid() % 200, ret), nprocs=2, join=True)
    ref = torch.randn(tables.global_layout.size, generator=torch.Generator().manual_seed(7))
    for rnd in range(3):
        combine_torch(tables, ref, [_client_arena(tables, l, i) + rnd for i, l in enumerate(levels)],
                      [tables.rates[l] for l in levels], masks)
    for r in range(2):
        got = torch.load(ret + ".%d" % r)
        assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5), (r, float((got - ref).abs().max()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_p2p_flag_protocol_stress(tmp_path):
    """release/acquire flag protocol under randomised per-rank delays, 24 back-to-back rounds with no host
    synchronisation in between, and a changing subset of surviving clients -- every replica still equals the oracle."""
    from hfl_b200.fed.federation import combine_torch
    tables, levels, masks, assignment = _agg_setup()
    ret = str(tmp_path / "ga.pt")
    rounds = 24
    mp.spawn(_agg_worker, args=(2, 29850 + os.getpid() % 100, ret, rounds, True), nprocs=2, join=True)
    ref = torch.randn(tables.global_layout.size, generator=torch.Generator().manual_seed(7))
    for rnd in range(rounds):
        act = _active(rnd, len(levels)) or list(range(len(levels)))
        combine_torch(tables, ref, [_client_arena(tables, levels[i], i) + rnd for i in act],
                      [tables.rates[levels[i]] for i in act], [masks[i] for i in act])
    for r in range(2):
        got = torch.load(ret + ".%d" % r)
        assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (r, float((got - ref).abs().max()))


def _worker(rank, world, port, backend, ret, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.parallel.dist import init_distributed, shutdown
    info = init_distributed()
    eng = Engine(_cfg(), info, options=EngineOptions(out_dir=out_dir, quiet=True, save=False, agg_backend=backend))
    eng.fit()
    if eng._peer is not None:
        eng._peer.check()
    if rank == 0:
        torch.save({"hist": dict(eng.logger.history)}, ret)
    shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("backend", ["p2p", "allreduce"])
def test_two_gpu_engine_runs(tmp_path, backend):
    """End-to-end 2-rank rounds (bf16, graphs, fused kernels): finite metrics, training loss falls."""
    import numpy as np
    ret = str(tmp_path / ("ret_%s.pt" % backend))
    mp.spawn(_worker, args=(2, 29700 + (os.getpid() % 200), backend, ret, str(tmp_path / "d")), nprocs=2, join=True)
    h = torch.load(ret, weights_only=False)["hist"]
    assert all(np.isfinite(v) for v in h["train/Local-Loss"] + h["test/Global-Loss"])
    assert h["train/Local-Loss"][-1] < h["train/Local-Loss"][0]


def _round_worker(rank, world, port, ret, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.parallel.dist import init_distributed, shutdown
    info = init_distributed()
    torch.manual_seed(0)
    eng = Engine(_cfg(), info, options=EngineOptions(out_dir=out_dir, quiet=True, save=False))
    eng.setup()
    w0 = eng.global_arena.float().cpu().clone()
    res = eng.run_round(1)
    torch.cuda.synchronize()
    if eng._peer is not None:
        eng._peer.check()
    if rank == 0:
        torch.save({"w0": w0, "w1": eng.global_arena.float().cpu().clone(), "train": res["train"], "test": res["test"],
                    "train_n": res["train_n"]}, ret)
    shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_rank_count_invariance(tmp_path):
    """The same communication round on 1 rank and on 2 ranks (clients sharded, P2P aggregation): same client sample, same
    number of training samples, and the global update W_g(1) - W_g(0) agrees in direction and size -- the clients' bf16
    trajectories differ only by kernel scheduling (atomics order, library algorithm choices), not by the rank count."""
    outs = []
    for world in (1, 2):
        ret = str(tmp_path / ("w%d.pt" % world))
        mp.spawn(_round_worker, args=(world, 29500 + world * 7 + (os.getpid() % 100), ret, str(tmp_path / ("d%d" % world))),
                 nprocs=world, join=True)
        outs.append(torch.load(ret, weights_only=False))
    a, b = outs
    assert torch.equal(a["w0"], b["w0"])
    assert a["train_n"] == b["train_n"]
    da, db = a["w1"] - a["w0"], b["w1"] - b["w0"]
    cos = float(torch.nn.functional.cosine_similarity(da, db, dim=0))
    ratio = float(db.norm() / da.norm())
    print("rank-count invariance: cos %.4f  norm ratio %.4f  loss %.4f / %.4f" % (
        cos, ratio, a["train"]["Local-Loss"], b["train"]["Local-Loss"]))
    assert cos > 0.9 and 0.9 < ratio < 1.1, (cos, ratio)
    assert abs(a["train"]["Local-Loss"] - b["train"]["Local-Loss"]) < 0.1 * max(1.0, abs(a["train"]["Local-Loss"]))
