import pytest
import torch

from hfl_b200.config.control import make_cfg
from hfl_b200.plan.spec import build_spec
from hfl_b200.utils.summary import format_summary, summarize


@pytest.mark.parametrize("data,model,control,extra,shape", [
    ("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1_bn_1_1", dict(classes_size=10), (3, 32, 32)),
    ("CIFAR10", "resnet18", "1_10_0.5_iid_fix_c1_bn_1_1", dict(classes_size=10), (3, 32, 32)),
    ("MNIST", "conv", "1_10_0.5_iid_fix_b1_bn_1_1", dict(classes_size=10), (1, 28, 28)),
])
def test_summary_matches_reference_hooks(ref, data, model, control, extra, shape):
    """Totals equal the reference's hook-based counter (incl. its per-parameter-tensor FLOP double count)."""
    m = ref.load()
    rcfg = ref.set_cfg(data, model, control, **extra)
    import importlib
    import os
    import sys
    cwd = os.getcwd()
    os.chdir("/root/reference/src")
    sys.path.insert(0, "/root/reference/src")
    argv = sys.argv
    sys.argv = ["summary.py", "--data_name", data, "--model_name", model, "--control_name", control, "--device", "cpu"]
    ref.reset_cfg()
    try:
        import tabulate  # noqa: F401
        sys.modules.pop("summary", None)
        rs = importlib.import_module("summary")
    finally:
        sys.argv = argv
        os.chdir(cwd)
        sys.path.remove("/root/reference/src")
    rcfg = ref.set_cfg(data, model, control, **extra)
    rmodel = getattr(m["models"], model)(model_rate=rcfg["global_model_rate"])
    loader = [{"img": [torch.randn(*shape)], "label": [torch.tensor(1)]}]
    rsum = rs.summarize(loader, rmodel)
    cfg = make_cfg(data, model, control, **extra)
    s = summarize(cfg)
    assert s["num_params"] == rsum["total_num_params"]
    assert s["num_flops"] == rsum["total_num_flops"]
    assert abs(s["space"] - rsum["total_space"]) < 1e-9
    assert "Total Number of Parameters" in format_summary(s)


def test_summary_levels_table():
    cfg = make_cfg("CIFAR10", "resnet18", "1_10_0.5_iid_fix_a1_bn_1_1", classes_size=10)
    want = {1.0: 11172170, 0.5: 2796714, 0.25: 701018, 0.125: 176178, 0.0625: 44510}
    spec = build_spec(cfg)
    for r, n in want.items():
        assert summarize(cfg, r, spec)["num_params"] == n
    assert abs(summarize(cfg, 1.0, spec)["num_macs_exact"] / 1e6 - 555.42) < 0.01
    cfg = make_cfg("WikiText2", "transformer", "1_10_0.5_iid_fix_a1_none_1_0", num_tokens=33278)
    assert summarize(cfg)["num_params"] == 19263486
