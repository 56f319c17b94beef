"""Local / Global test metrics == the reference's per-user-loader test() (same weights, same running statistics)."""
import numpy as np
import pytest
import torch

from hfl_b200.config.control import make_cfg
from hfl_b200.data.datasets import synth_vision
from hfl_b200.fed.engine import Engine, EngineOptions


class _RefDS(torch.utils.data.Dataset):
    """What the reference's dataset classes yield after their transform: normalised CHW float + label."""

    def __init__(self, ds):
        mean = torch.tensor(ds.mean).view(-1, 1, 1)
        std = torch.tensor(ds.std).view(-1, 1, 1)
        self.x = (ds.img.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
        self.target = ds.target.tolist()

    def __len__(self):
        return len(self.target)

    def __getitem__(self, i):
        return {"img": self.x[i], "label": torch.tensor(self.target[i])}


@pytest.mark.parametrize("mask", [1, 0])
def test_local_global_metrics_match_reference(ref, tmp_path, mask):
    m = ref.load()
    control = "1_10_0.5_non-iid-2_fix_a1-c1_bn_1_%d" % mask
    cfg = make_cfg("MNIST", "conv", control, synthetic=True, synthetic_size={"train": 400, "test": 200})
    cfg["num_epochs"] = {"global": 1, "local": 1}
    eng = Engine(cfg, options=EngineOptions(out_dir=str(tmp_path), quiet=True, save=False))
    res = eng.run_round(1)
    # reference side: track=True model with our weights + our running statistics, its own test()
    rcfg = ref.set_cfg("MNIST", "conv", control, classes_size=10)
    import importlib
    import os
    import sys
    cwd = os.getcwd()
    os.chdir("/root/reference/src")
    sys.path.insert(0, "/root/reference/src")
    argv = sys.argv
    sys.argv = ["train_classifier_fed.py", "--data_name", "MNIST", "--model_name", "conv", "--control_name", control,
                "--device", "cpu"]
    ref.reset_cfg()
    try:
        sys.modules.pop("train_classifier_fed", None)
        T = importlib.import_module("train_classifier_fed")
    finally:
        sys.argv = argv
        os.chdir(cwd)
        sys.path.remove("/root/reference/src")
    rcfg = ref.set_cfg("MNIST", "conv", control, classes_size=10)
    rcfg["model_tag"] = "t"
    rcfg["metric_name"] = {"train": {"Local": ["Local-Loss", "Local-Accuracy"]},
                           "test": {"Local": ["Local-Loss", "Local-Accuracy"], "Global": ["Global-Loss", "Global-Accuracy"]}}
    rmodel = m["models"].conv(model_rate=1.0, track=True)
    sd = eng.state_dict()
    for name, (rm, rv) in eng.running.items():
        sd[name + ".running_mean"], sd[name + ".running_var"] = rm.clone(), rv.clone()
        sd[name + ".num_batches_tracked"] = torch.tensor(1)
    rmodel.load_state_dict(sd)
    rlogger = T.Logger(None)
    rlogger.safe = lambda w: None
    import builtins
    T.test(_RefDS(eng.dataset["test"]), eng.data_split["test"], eng.label_split, rmodel, rlogger, 1)
    for k in ("Local-Loss", "Local-Accuracy", "Global-Loss", "Global-Accuracy"):
        assert res["test"][k] == pytest.approx(rlogger.mean["test/" + k], rel=2e-4, abs=2e-4), k
