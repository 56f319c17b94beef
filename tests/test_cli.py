"""CLI parity: scripts run end to end on CPU with synthetic data and write the reference's output layout."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args, cwd):
    cmd = [sys.executable, os.path.join(ROOT, script)] + list(args)
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return p.stdout


def test_fed_train_then_test_scripts(tmp_path, monkeypatch):
    import hfl_b200.data.datasets as D
    env_sizes = "HFL_SYNTH_SIZES"
    os.environ[env_sizes] = "800,200"
    try:
        control = "1_10_0.3_iid_fix_a1-e1_bn_1_1"
        out = _run("train_classifier_fed.py", "--data_name", "MNIST", "--model_name", "conv", "--control_name", control,
                   "--device", "cpu", "--rounds", "2", "--synthetic", "1", cwd=str(tmp_path))
        assert "Train Epoch: 2" in out and "Global-Accuracy" in out
        tag = "0_MNIST_label_conv_" + control
        assert os.path.exists(tmp_path / "output" / "model" / (tag + "_checkpoint.pt"))
        assert os.path.exists(tmp_path / "output" / "model" / (tag + "_best.pt"))
        out = _run("test_classifier_fed.py", "--data_name", "MNIST", "--model_name", "conv", "--control_name", control,
                   "--device", "cpu", "--synthetic", "1", cwd=str(tmp_path))
        from hfl_b200.ckpt.compat import load
        res = load(str(tmp_path / "output" / "result" / (tag + ".pt")))
        assert set(res) == {"cfg", "epoch", "logger"} and set(res["logger"]) == {"train", "test"}
        assert "test/Global-Accuracy" in res["logger"]["test"].history
    finally:
        os.environ.pop(env_sizes, None)


def test_central_train_script(tmp_path):
    os.environ["HFL_SYNTH_SIZES"] = "600,200"
    try:
        out = _run("train_classifier.py", "--data_name", "MNIST", "--model_name", "conv", "--control_name",
                   "0_1_1_none_fix_c1_bn_1_1", "--device", "cpu", "--rounds", "2", "--synthetic", "1", cwd=str(tmp_path))
        assert "Train Epoch: 2" in out and "Accuracy" in out
    finally:
        os.environ.pop("HFL_SYNTH_SIZES", None)


def test_supervisor_resumes_after_crash(tmp_path):
    """a rank dying mid-run -> supervisor relaunches with --resume_mode 1 from the per-round checkpoint."""
    env = dict(os.environ, HFL_SYNTH_SIZES="600,200", HFL_FAULT_CRASH_ROUND="2")
    control = "1_10_0.3_iid_fix_a1-e1_bn_1_1"
    cmd = [sys.executable, os.path.join(ROOT, "supervise.py"), "--max_restarts", "2", "--backoff", "0.1", "--",
           sys.executable, os.path.join(ROOT, "train_classifier_fed.py"), "--data_name", "MNIST", "--model_name", "conv",
           "--control_name", control, "--device", "cpu", "--rounds", "3", "--synthetic", "1"]
    p = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "attempt 1 exited with 17" in p.stderr and "Resume from 2" in p.stdout
    from hfl_b200.ckpt.compat import load
    ck = load(str(tmp_path / "output" / "model" / ("0_MNIST_label_conv_" + control + "_checkpoint.pt")))
    assert ck["epoch"] == 5           # --rounds counts from the resumed epoch: 1 (crash before 2) then 2, 3, 4


def test_fault_drop_changes_only_aggregation(tmp_path):
    """Fault injection: lost clients do not contribute; with p=0 the round is bit-identical to the default."""
    from hfl_b200.config.control import make_cfg
    from hfl_b200.fed.engine import Engine, EngineOptions

    def run(p):
        cfg = make_cfg("MNIST", "conv", "1_10_0.5_iid_fix_a1-e1_bn_1_1", synthetic=True, synthetic_size={"train": 400, "test": 100})
        eng = Engine(cfg, options=EngineOptions(out_dir=str(tmp_path), quiet=True, save=False, fault_drop=p))
        eng.setup()
        res = eng.run_round()
        return res, eng.global_arena.clone()

    r0, a0 = run(0.0)
    r1, a1 = run(0.6)
    assert r0["lost"] == [] and len(r1["lost"]) >= 1 and len(r1["lost"]) < r1["num_active"]
    assert r0["train"] == r1["train"]                     # local training is unaffected
    assert not torch.equal(a0, a1)
