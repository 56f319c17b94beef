"""Property tests of the heterogeneous-FedAvg algebra (hypothesis, CPU): the invariants SURVEY section 4 lists --
coverage count = number of covering clients, uncovered coordinates unchanged, homogeneous case == plain FedAvg,
distribute-then-combine of untouched clients is the identity."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from hfl_b200.config.control import make_cfg
from hfl_b200.fed.federation import combine_torch, distribute_torch
from hfl_b200.plan.layout import build_tables, index_maps
from hfl_b200.plan.spec import build_spec

_CFG = make_cfg("MNIST", "conv", "1_10_0.5_iid_fix_a1-b1-c1-d1-e1_bn_1_1", classes_size=10)
_SPEC = build_spec(_CFG)
_RATES = sorted(set(_CFG["model_rate"]), reverse=True)
_TABLES = build_tables(_SPEC, 1.0, _RATES, channels_last=False)
_G = _TABLES.global_layout.size

levels_st = st.lists(st.integers(min_value=0, max_value=len(_RATES) - 1), min_size=1, max_size=6)
mask_st = st.lists(st.booleans(), min_size=10, max_size=10).filter(lambda m: any(m))


def _masks(draws, n):
    return [torch.tensor(draws[i % len(draws)], dtype=torch.uint8) for i in range(n)]


@settings(max_examples=25, deadline=None)
@given(levels=levels_st, masks=st.lists(mask_st, min_size=1, max_size=6), seed=st.integers(0, 2 ** 16))
def test_combine_is_mean_over_covering_clients(levels, masks, seed):
    """Clients filled with a constant v_i: every global coordinate ends up as the mean of v_i over exactly the clients
    whose slice (and, for label rows, whose label mask) covers it; coordinates nobody covers keep their old value."""
    g = torch.Generator().manual_seed(seed)
    old = torch.randn(_G, generator=g)
    arena = old.clone()
    rates = [_RATES[l] for l in levels]
    lm = _masks(masks, len(levels))
    vals = [float(i + 1) for i in range(len(levels))]
    clients = [torch.full((_TABLES.level_layouts[l].size,), v) for l, v in zip(levels, vals)]
    combine_torch(_TABLES, arena, clients, rates, lm)
    gv, ov = _TABLES.global_layout.views(arena), _TABLES.global_layout.views(old)
    for p in _SPEC.params:
        acc = torch.zeros(gv[p.name].shape)
        cnt = torch.zeros(gv[p.name].shape)
        for rate, v, m in zip(rates, vals, lm):
            idx = index_maps(_SPEC, rate, 1.0)[p.name]
            if p.inp is None:
                oi = idx
                if p.label_rows:
                    oi = oi[m[oi].bool()]
                acc[oi] += v
                cnt[oi] += 1
            else:
                oi, ii = idx
                if p.label_rows:
                    oi = oi[m[oi].bool()]
                acc[oi[:, None], ii[None, :]] += v
                cnt[oi[:, None], ii[None, :]] += 1
        want = torch.where(cnt > 0, acc / cnt.clamp(min=1), ov[p.name])
        assert torch.allclose(gv[p.name], want, atol=1e-6), p.name


@settings(max_examples=15, deadline=None)
@given(n=st.integers(1, 6), seed=st.integers(0, 2 ** 16))
def test_homogeneous_case_is_plain_fedavg(n, seed):
    g = torch.Generator().manual_seed(seed)
    arena = torch.randn(_G, generator=g)
    clients = [torch.randn(_G, generator=g) for _ in range(n)]
    ones = [torch.ones(10, dtype=torch.uint8)] * n
    combine_torch(_TABLES, arena, clients, [1.0] * n, ones)
    mean = torch.stack(clients).mean(0)
    got, want = _TABLES.global_layout.views(arena), _TABLES.global_layout.views(mean)
    for k in got:                      # (alignment padding between tensors is not a parameter and is never touched)
        assert torch.allclose(got[k], want[k], atol=1e-6), k


@settings(max_examples=15, deadline=None)
@given(levels=levels_st, masks=st.lists(mask_st, min_size=1, max_size=6), seed=st.integers(0, 2 ** 16))
def test_distribute_then_combine_is_identity(levels, masks, seed):
    """Slicing W_g into any set of sub-networks and averaging them back without training leaves W_g unchanged."""
    g = torch.Generator().manual_seed(seed)
    arena = torch.randn(_G, generator=g)
    before = arena.clone()
    rates = [_RATES[l] for l in levels]
    clients = [torch.zeros(_TABLES.level_layouts[l].size) for l in levels]
    distribute_torch(_TABLES, arena, clients, rates)
    combine_torch(_TABLES, arena, clients, rates, _masks(masks, len(levels)))
    assert torch.allclose(arena, before, atol=1e-6)


@settings(max_examples=10, deadline=None)
@given(level=st.integers(1, len(_RATES) - 1), seed=st.integers(0, 2 ** 16))
def test_narrow_clients_leave_the_wide_remainder_alone(level, seed):
    g = torch.Generator().manual_seed(seed)
    arena = torch.randn(_G, generator=g)
    before = arena.clone()
    rate = _RATES[level]
    client = torch.randn(_TABLES.level_layouts[level].size, generator=g)
    combine_torch(_TABLES, arena, [client], [rate], [torch.ones(10, dtype=torch.uint8)])
    changed = (arena != before)
    covered = torch.zeros(_G, dtype=torch.bool)
    cv = _TABLES.global_layout.views(covered)
    for p in _SPEC.params:
        idx = index_maps(_SPEC, rate, 1.0)[p.name]
        if p.inp is None:
            cv[p.name][idx] = True
        else:
            cv[p.name][idx[0][:, None], idx[1][None, :]] = True
    assert not bool((changed & ~covered).any())
    assert int(covered.sum()) == sum(e.numel for e in _TABLES.level_layouts[level].entries.values())
