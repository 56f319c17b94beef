"""Logger / metrics / optimizer+scheduler factories / control parsing vs the reference implementations."""
import importlib
import os
import sys

import pytest
import torch

from hfl_b200.config.control import make_cfg, parse_control_name
from hfl_b200.utils.logger import Logger
from hfl_b200.utils.metrics import Metric
from hfl_b200.utils.optim import make_optimizer, make_scheduler


def _ref_module(name):
    cwd = os.getcwd()
    os.chdir("/root/reference/src")
    sys.path.insert(0, "/root/reference/src")
    try:
        return importlib.import_module(name)
    finally:
        os.chdir(cwd)
        sys.path.remove("/root/reference/src")


def test_logger_matches_reference(ref, tmp_path, capsys):
    ref.load()
    rlog_mod = _ref_module("logger")
    sys.modules.pop("logger", None)
    a, b = rlog_mod.Logger(None), Logger(None, tensorboard=False)
    seq = [({"Local-Loss": 1.0, "Local-Accuracy": 50.0}, 10), ({"Local-Loss": 0.5, "Local-Accuracy": 70.0}, 30),
           ({"Local-Loss": 0.25, "Local-Accuracy": [80.0, 90.0]}, 5)]
    for lg in (a, b):
        for res, n in seq[:2]:
            lg.append(res, "train", n=n)
        lg.append({"V": [1.0, 2.0]}, "train", n=4)
        lg.append({"V": [3.0, 4.0]}, "train", n=(4, 12))
        lg.append({"info": ["Model: x", "Train Epoch: 1(0%)", "tail"]}, "train", mean=False)
    a.write("train", ["Local-Loss", "Local-Accuracy"])
    out_a = capsys.readouterr().out
    b.write("train", ["Local-Loss", "Local-Accuracy"])
    out_b = capsys.readouterr().out
    assert out_a == out_b
    assert dict(a.mean) == dict(b.mean) and dict(a.counter) == dict(b.counter)
    a.safe(False), b.safe(False)
    assert dict(a.history) == dict(b.history)
    a.reset(), b.reset()
    assert dict(a.mean) == dict(b.mean) == {}


def test_metrics_match_reference(ref):
    ref.load()
    rmet = _ref_module("metrics")
    torch.manual_seed(0)
    score = torch.randn(7, 10)
    label = torch.randint(0, 10, (7,))
    out = {"loss": torch.nn.functional.cross_entropy(score, label), "score": score}
    inp = {"label": label}
    names = ["Loss", "Accuracy", "Local-Loss", "Local-Accuracy", "Global-Loss", "Global-Accuracy"]
    assert Metric().evaluate(names, inp, out) == pytest.approx(rmet.Metric().evaluate(names, inp, out))
    lm_score = torch.randn(2, 30, 16)
    lm_label = torch.randint(0, 30, (2, 16))
    o2 = {"loss": torch.nn.functional.cross_entropy(lm_score, lm_label), "score": lm_score}
    for n in ("Perplexity", "Local-Perplexity", "Global-Perplexity"):
        assert Metric().evaluate([n], {"label": lm_label}, o2)[n] == pytest.approx(
            rmet.Metric().evaluate([n], {"label": lm_label}, o2)[n], rel=1e-6)


@pytest.mark.parametrize("opt", ["SGD", "RMSprop", "Adam", "Adamax"])
@pytest.mark.parametrize("sched", ["None", "StepLR", "MultiStepLR", "ExponentialLR", "CosineAnnealingLR", "CyclicLR"])
def test_optimizer_scheduler_factories(ref, opt, sched):
    m = ref.load()
    rcfg = ref.set_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1_bn_1_1", classes_size=10)
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1_bn_1_1", classes_size=10)
    for c in (rcfg, cfg):
        c["optimizer_name"], c["scheduler_name"] = opt, sched
        c["step_size"], c["min_lr"] = 2, 1e-4
    if sched == "CyclicLR" and opt in ("Adam", "Adamax"):
        pytest.skip("CyclicLR needs a momentum optimizer")
    lin_a, lin_b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    oa = m["utils"].make_optimizer(lin_a, rcfg["lr"])
    ob = make_optimizer(list(lin_b.parameters()), cfg, cfg["lr"])
    assert type(oa) is type(ob) and oa.defaults == ob.defaults
    sa, sb = m["utils"].make_scheduler(oa), make_scheduler(ob, cfg)
    lrs_a, lrs_b = [], []
    for _ in range(6):
        oa.step(), ob.step()
        sa.step(), sb.step()
        lrs_a.append(oa.param_groups[0]["lr"])
        lrs_b.append(ob.param_groups[0]["lr"])
    assert lrs_a == pytest.approx(lrs_b)


def test_control_name_grammar():
    c = parse_control_name("1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1")
    assert list(c.keys()) == ["fed", "num_users", "frac", "data_split_mode", "model_split_mode", "model_mode", "norm",
                              "scale", "mask"]
    with pytest.raises(ValueError):
        parse_control_name("1_100_0.1_iid")
    cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_fix_a2-b8_bn_1_1")
    assert cfg["model_rate"].count(1.0) == 20 and cfg["model_rate"].count(0.5) == 80 and cfg["non-iid-n"] == 2
    cfg = make_cfg("CIFAR10", "resnet18", "1_7_1.0_iid_fix_a1-b2_gn_0_0")      # 7 users, proportion sum 3 -> tail padded
    assert cfg["model_rate"] == [1.0, 1.0, 0.5, 0.5, 0.5, 0.5, 0.5] and not cfg["scale"] and not cfg["mask"]
    with pytest.raises(ValueError):
        make_cfg("CIFAR10", "resnet18", "1_10_1.0_iid_fix_a1_xx_1_1")
    with pytest.raises(ValueError):
        make_cfg("WikiText2", "transformer", "1_10_1.0_non-iid-2_fix_a1_none_1_1")


@pytest.mark.parametrize("control", ["1_100_0.1_iid_fix_a2-b8_bn_1_1", "1_100_0.1_non-iid-2_dynamic_a1-c3-e6_in_0_1",
                                     "0_1_1_none_fix_c1_none_1_0", "1_7_0.5_iid_fix_a1-b2_ln_1_1"])
@pytest.mark.parametrize("data,model", [("MNIST", "conv"), ("CIFAR10", "resnet18")])
def test_process_control_matches_reference(ref, data, model, control):
    rcfg = ref.set_cfg(data, model, control)
    cfg = make_cfg(data, model, control)
    for k in ("fed", "num_users", "frac", "data_split_mode", "model_split_mode", "model_mode", "norm", "scale", "mask",
              "global_model_mode", "global_model_rate", "model_rate", "data_shape", "optimizer_name", "lr", "momentum",
              "weight_decay", "scheduler_name", "factor", "num_epochs", "batch_size", "milestones"):
        assert cfg[k] == rcfg[k], k
    if control.split("_")[4] == "dynamic":
        assert cfg["proportion"] == pytest.approx(rcfg["proportion"])


def test_aggregate_traffic_summary():
    from hfl_b200.utils.commstats import aggregate_traffic
    r = aggregate_traffic([1000, 250], [0, 1, 1], 1000, 1, 0.5, hbm_gbs=6478.0)
    assert r["client_bytes"] == 4 * 1500 and r["publish_bytes"] == 4000 and "hbm_roofline_frac" in r
    r8 = aggregate_traffic([1000, 250], [0, 1, 1], 1000, 8, 1.0)
    assert r8["publish_bytes"] == 32000 and 0 < r8["nvlink_roofline_frac"] < 1
    assert "gbs" not in aggregate_traffic([10], [0], 10, 2, 0.0)
