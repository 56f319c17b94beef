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
