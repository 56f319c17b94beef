"""Engine integration on CPU: round loop, gloo world_size=2 == single process, checkpoint/resume round trip."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from hfl_b200.config.control import make_cfg
from hfl_b200.fed.engine import Engine, EngineOptions

CONTROL = "1_8_1.0_iid_fix_a1_bn_1_1"          # BASELINE.json config 1 (plumbing)


def _cfg(control=CONTROL, rounds=2, **kw):
    cfg = make_cfg("MNIST", "conv", control, synthetic=True, synthetic_size={"train": 640, "test": 160}, **kw)
    cfg["num_epochs"] = {"global": rounds, "local": 1}
    return cfg


def _run(rank, world, port, out_dir, control, ret, fault_drop=0.0):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hfl_b200.parallel.dist import init_distributed, shutdown
    info = init_distributed("cpu")
    eng = Engine(_cfg(control), info, options=EngineOptions(out_dir=out_dir, quiet=True, save=(rank == 0),
                                                           fault_drop=fault_drop))
    eng.fit()
    if rank == 0:
        torch.save({"arena": eng.global_arena.clone(), "hist": dict(eng.logger.history)}, ret)
    shutdown()


@pytest.mark.parametrize("control", [CONTROL, "1_10_0.5_non-iid-2_dynamic_a1-c1-e1_bn_1_1"])
def test_gloo_two_ranks_match_single(tmp_path, control):
    torch.manual_seed(0)
    single = Engine(_cfg(control), options=EngineOptions(out_dir=str(tmp_path / "s"), quiet=True))
    single.fit()
    ret = str(tmp_path / "ret.pt")
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_run, args=(2, port, str(tmp_path / "d"), control, ret), nprocs=2, join=True)
    got = torch.load(ret, weights_only=False)
    assert torch.allclose(got["arena"], single.global_arena, atol=1e-5, rtol=1e-4)
    for k, v in single.logger.history.items():
        assert got["hist"][k] == pytest.approx(v, rel=1e-4, abs=1e-5), k


def test_fault_injection_is_rank_consistent(tmp_path):
    """Lost clients are drawn identically on every rank: two gloo ranks with fault_drop reproduce the single-process
    run with the same drop probability (the shared-seed contract that replaces a broadcast of the survivor list)."""
    control = "1_10_0.5_non-iid-2_dynamic_a1-c1-e1_bn_1_1"
    single = Engine(_cfg(control), options=EngineOptions(out_dir=str(tmp_path / "s"), quiet=True, fault_drop=0.4))
    single.fit()
    ret = str(tmp_path / "ret.pt")
    port = 29910 + (os.getpid() % 80)
    mp.spawn(_run, args=(2, port, str(tmp_path / "d"), control, ret, 0.4), nprocs=2, join=True)
    got = torch.load(ret, weights_only=False)
    assert torch.allclose(got["arena"], single.global_arena, atol=1e-5, rtol=1e-4)


def test_supervisor_resume_flag_rewrite():
    from hfl_b200.cli.supervise import with_resume
    assert with_resume(["python", "t.py", "--x", "1"]) == ["python", "t.py", "--x", "1", "--resume_mode", "1"]
    assert with_resume(["t.py", "--resume_mode", "0", "--y"]) == ["t.py", "--y", "--resume_mode", "1"]
    assert with_resume(["t.py", "--resume_mode=2"]) == ["t.py", "--resume_mode", "1"]


def test_checkpoint_resume_roundtrip(tmp_path):
    out = str(tmp_path / "o")
    e1 = Engine(_cfg(rounds=3), options=EngineOptions(out_dir=out, quiet=True))
    e1.fit(num_rounds=2)
    tag = e1.cfg["model_tag"]
    assert os.path.exists(os.path.join(out, "model", tag + "_checkpoint.pt"))
    assert os.path.exists(os.path.join(out, "model", tag + "_best.pt"))
    e2 = Engine(_cfg(rounds=3, resume_mode=1), options=EngineOptions(out_dir=out, quiet=True))
    e2.setup()
    assert e2.epoch == 3
    assert torch.allclose(e2.global_arena, e1.global_arena)
    assert e2.data_split == e1.data_split and e2.label_split == e1.label_split
    assert len(e2.logger.history["train/Local-Loss"]) == 2
    e2.fit()
    assert e2.epoch == 4 and len(e2.logger.history["test/Global-Accuracy"]) == 3
    # resume_mode 2: model + splits only, restart at epoch 1
    e3 = Engine(_cfg(rounds=3, resume_mode=2), options=EngineOptions(out_dir=out, quiet=True))
    e3.setup()
    assert e3.epoch == 1 and torch.allclose(e3.global_arena, e2.global_arena)


def test_checkpoint_is_reference_compatible(ref, tmp_path):
    """A checkpoint written here loads through the reference's own load()/resume machinery and vice versa."""
    m = ref.load()
    out = str(tmp_path / "o")
    e1 = Engine(_cfg(rounds=1), options=EngineOptions(out_dir=out, quiet=True))
    e1.fit()
    path = os.path.join(out, "model", e1.cfg["model_tag"] + "_checkpoint.pt")
    # reference-side unpickle needs its own `logger` module importable
    sys.path.insert(0, "/root/reference/src")
    cwd = os.getcwd()
    os.chdir("/root/reference/src")
    try:
        import logger as rlogger  # noqa: F401
        ck = m["utils"].load(path) if False else torch.load(path, map_location="cpu", weights_only=False)
    finally:
        os.chdir(cwd)
        sys.path.remove("/root/reference/src")
        sys.modules.pop("logger", None)
    assert set(ck.keys()) == {"cfg", "epoch", "data_split", "label_split", "model_dict", "optimizer_dict",
                              "scheduler_dict", "logger"}
    assert type(ck["logger"]).__module__ == "logger" and ck["epoch"] == 2
    rcfg = ref.set_cfg("MNIST", "conv", CONTROL, classes_size=10)
    rmodel = m["models"].conv(model_rate=1)
    rmodel.load_state_dict(ck["model_dict"])          # same keys / shapes as the reference module
    ropt = m["utils"].make_optimizer(rmodel, rcfg["lr"])
    ropt.load_state_dict(ck["optimizer_dict"])
    rsch = m["utils"].make_scheduler(ropt)
    rsch.load_state_dict(ck["scheduler_dict"])
    # and the other direction: a reference-written checkpoint resumes here
    rlog = None
    sys.path.insert(0, "/root/reference/src")
    os.chdir("/root/reference/src")
    try:
        import logger as rlogger
        rlog = rlogger.Logger(str(tmp_path / "runs"))
        rlog.append({"Local-Loss": 1.5}, "train", n=10)
        rlog.safe(False)
        obj = {"cfg": dict(rcfg), "epoch": 5, "data_split": e1.data_split, "label_split": e1.label_split,
               "model_dict": rmodel.state_dict(), "optimizer_dict": ropt.state_dict(),
               "scheduler_dict": rsch.state_dict(), "logger": rlog}
        os.makedirs(os.path.join(out, "model"), exist_ok=True)
        torch.save(obj, os.path.join(out, "model", e1.cfg["model_tag"] + "_checkpoint.pt"), pickle_protocol=2)
    finally:
        os.chdir(cwd)
        sys.path.remove("/root/reference/src")
        sys.modules.pop("logger", None)
    e2 = Engine(_cfg(rounds=6, resume_mode=1), options=EngineOptions(out_dir=out, quiet=True))
    e2.setup()
    assert e2.epoch == 5 and e2.logger.history["train/Local-Loss"] == [1.5]


def test_async_eval_policy_and_late_test_logging(tmp_path):
    """Asynchronous evaluation is a CUDA-only mode (automatic from 4 ranks up, vision models): on CPU the engine never
    selects it, and `_log_test` -- the path fit() uses for the one-round-late metrics of that mode -- feeds the logger
    exactly like the in-round evaluation."""
    assert EngineOptions().async_eval is None
    eng = Engine(_cfg(rounds=1), options=EngineOptions(out_dir=str(tmp_path / "a"), quiet=True, save=False,
                                                      async_eval=True))
    eng.setup()
    assert not eng._async_eval_ok()                       # no CUDA device: forced on or not, the mode stays off
    res = eng.run_round(1)
    assert "test" in res and "prev_eval" not in res and "eval_pending" not in res
    eng.logger.safe(True)
    eng._log_test({"epoch": 1, "test": res["test"], "_test_n": res.get("_test_n", {})})
    eng.logger.safe(False)
    key = "test/Global-Accuracy"
    assert key in eng.logger.history and eng.logger.history[key][-1] == pytest.approx(res["test"]["Global-Accuracy"])
