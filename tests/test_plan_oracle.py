"""Slice planner / distribute / combine vs the reference Federation (oracle, CPU)."""
import copy

import numpy as np
import pytest
import torch

from hfl_b200.config.control import make_cfg
from hfl_b200.fed.federation import Federation
from hfl_b200.plan.layout import Layout, index_maps
from hfl_b200.plan.spec import build_spec

CASES = [
    ("MNIST", "conv", "1_10_0.5_iid_fix_a1-b1-c1-d1-e1_bn_1_1", dict(classes_size=10)),
    ("MNIST", "conv", "1_10_0.5_iid_fix_b1-e1_none_1_1", dict(classes_size=10)),
    ("CIFAR10", "resnet18", "1_10_0.6_non-iid-2_fix_a1-b1-c1-d1-e1_bn_1_1", dict(classes_size=10)),
    ("CIFAR10", "resnet18", "1_10_0.5_iid_dynamic_a1-c2-e1_gn_0_0", dict(classes_size=10)),
    ("WikiText2", "transformer", "1_10_0.5_iid_fix_a1-b1-c1-d1-e1_none_1_0", dict(num_tokens=97)),
    ("WikiText2", "transformer", "1_10_0.5_iid_fix_b5-d5_none_1_1", dict(num_tokens=50)),
]


def _label_split(cfg, n_rows, seed):
    g = np.random.RandomState(seed)
    return {u: sorted(g.choice(n_rows, size=max(2, n_rows // 3), replace=False).tolist())
            for u in range(cfg["num_users"])}


@pytest.mark.parametrize("data_name,model_name,control,extra", CASES)
def test_matches_reference(ref, data_name, model_name, control, extra):
    m = ref.load()
    rcfg = ref.set_cfg(data_name, model_name, control, **extra)
    cfg = make_cfg(data_name, model_name, control, **extra)
    for k in ("model_rate", "global_model_rate", "num_users", "frac", "norm", "scale", "mask"):
        assert rcfg[k] == cfg[k], k
    torch.manual_seed(0)
    if model_name == "transformer":
        rmodel = m["models"].transformer(model_rate=rcfg["global_model_rate"])
    else:
        rmodel = getattr(m["models"], model_name)(model_rate=rcfg["global_model_rate"])
    gsd = rmodel.state_dict()
    spec = build_spec(cfg)
    # key names, order and shapes
    glay = Layout(spec, cfg["global_model_rate"])
    assert list(gsd.keys()) == list(glay.entries.keys())
    for k, v in gsd.items():
        assert tuple(v.shape) == glay.entries[k].shape, k
    n_rows = cfg.get("classes_size", cfg.get("num_tokens"))
    label_split = _label_split(cfg, n_rows, 1)

    arena = torch.zeros(glay.size)
    glay.load_state_dict(arena, gsd)
    torch.manual_seed(123)
    rfederation = m["fed"].Federation(copy.deepcopy(gsd), rcfg["model_rate"], label_split)
    torch.manual_seed(123)
    federation = Federation(cfg, spec, arena, label_split, backend="torch")
    assert np.allclose(rfederation.model_rate, federation.model_rate)

    user_idx = [7, 2, 5, 0, 9]
    torch.manual_seed(5)
    rlocal, rparam_idx = rfederation.distribute(user_idx)
    torch.manual_seed(5)
    arenas = federation.distribute(user_idx)
    assert np.allclose(rfederation.model_rate, federation.model_rate)
    for mi, u in enumerate(user_idx):
        rate = float(federation.model_rate[u])
        maps = index_maps(spec, rate, cfg["global_model_rate"])
        lay = federation.tables.level_layouts[federation.tables.level_id(rate)]
        views = lay.views(arenas[mi])
        for k in gsd:
            ridx = rparam_idx[mi][k]
            if isinstance(ridx, tuple):
                assert torch.equal(ridx[0], maps[k][0]) and torch.equal(ridx[1], maps[k][1]), k
            else:
                assert torch.equal(ridx, maps[k]), k
            assert torch.equal(rlocal[mi][k], views[k]), k

    # fake local training: perturb every local tensor deterministically, identically on both sides
    g = torch.Generator().manual_seed(9)
    for mi in range(len(user_idx)):
        for k in gsd:
            noise = torch.randn(rlocal[mi][k].shape, generator=g)
            rlocal[mi][k] = rlocal[mi][k] + noise
            lay = federation.tables.level_layouts[federation.tables.level_id(float(federation.model_rate[user_idx[mi]]))]
            lay.views(arenas[mi])[k].add_(noise)
    rfederation.combine(rlocal, rparam_idx, user_idx)
    federation.combine(arenas, user_idx)
    new = glay.views(arena)
    for k, v in rfederation.global_parameters.items():
        assert torch.allclose(v, new[k], atol=1e-6, rtol=1e-6), k
