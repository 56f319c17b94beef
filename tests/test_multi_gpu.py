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


def _worker(rank, world, port, backend, ret, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.parallel.dist import init_distributed, shutdown
    info = init_distributed()
    # fp32 compute + no graphs: results must not depend on which rank trained which client
    eng = Engine(_cfg(), info, options=EngineOptions(out_dir=out_dir, quiet=True, save=False, agg_backend=backend,
                                                     compute_dtype="fp32", use_graph=False))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    eng.fit()
    if eng._peer is not None:
        eng._peer.check()
    if rank == 0:
        torch.save({"arena": eng.global_arena.float().cpu().clone(), "hist": dict(eng.logger.history)}, ret)
    shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("backend", ["p2p", "allreduce"])
def test_two_gpu_aggregation_matches_single(tmp_path, backend):
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.parallel.dist import DistInfo
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    dev = torch.device("cuda", 0)
    single = Engine(_cfg(), DistInfo(0, 1, 0, dev, None),
                    options=EngineOptions(out_dir=str(tmp_path / "s"), quiet=True, save=False, compute_dtype="fp32",
                                          use_graph=False))
    single.fit()
    ret = str(tmp_path / ("ret_%s.pt" % backend))
    port = 29700 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, backend, ret, str(tmp_path / "d")), nprocs=2, join=True)
    got = torch.load(ret, weights_only=False)
    ref = single.global_arena.float().cpu()
    assert torch.allclose(got["arena"], ref, atol=2e-3, rtol=2e-2), float((got["arena"] - ref).abs().max())
    for k in ("train/Local-Loss", "test/Global-Loss"):
        assert got["hist"][k] == pytest.approx(single.logger.history[k], rel=5e-2), k
