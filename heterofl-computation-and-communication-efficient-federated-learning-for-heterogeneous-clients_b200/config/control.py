"""Configuration + ``control_name`` grammar.

Parity targets (reference, read-only): default config ``src/config.yml:1-55``; control parsing in the entry
points ``src/train_classifier_fed.py:20-34``; the hyper-parameter table ``src/utils.py:113-215``
(``process_control``).  The reference keeps one global mutable dict; here a config is an explicit plain ``dict``
(so it pickles into checkpoints under the same keys) created by :func:`make_cfg` and never a module global.

control_name = ``fed_numusers_frac_datasplit_modelsplit_modelmode_norm_scale_mask``
  e.g. ``1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1``.
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict
from typing import Any, Dict, List

CONTROL_KEYS = ("fed", "num_users", "frac", "data_split_mode", "model_split_mode", "model_mode", "norm", "scale",
                "mask")

#: complexity level -> width multiplier (reference ``src/utils.py:114``)
MODEL_SPLIT_RATE = {"a": 1.0, "b": 0.5, "c": 0.25, "d": 0.125, "e": 0.0625}

VISION_DATA = ("MNIST", "CIFAR10", "CIFAR100", "EMNIST", "FashionMNIST")
LM_DATA = ("PennTreebank", "WikiText2", "WikiText103")

DEFAULT_CFG: Dict[str, Any] = OrderedDict([
    ("control", OrderedDict([("fed", "1"), ("num_users", "100"), ("frac", "0.1"), ("data_split_mode", "iid"),
                             ("model_split_mode", "fix"), ("model_mode", "a1"), ("norm", "bn"), ("scale", "1"),
                             ("mask", "1")])),
    ("data_name", "CIFAR10"),
    ("subset", "label"),
    ("batch_size", {"train": 128, "test": 128}),
    ("shuffle", {"train": True, "test": False}),
    ("num_workers", 0),
    ("model_name", "resnet18"),
    ("metric_name", {"train": ["Loss", "Accuracy"], "test": ["Loss", "Accuracy"]}),
    ("optimizer_name", "Adam"),
    ("lr", 3.0e-4),
    ("momentum", 0.9),
    ("weight_decay", 5.0e-4),
    ("scheduler_name", "None"),
    ("step_size", 1),
    ("milestones", [100, 150]),
    ("patience", 10),
    ("threshold", 1.0e-3),
    ("factor", 0.5),
    ("min_lr", 1.0e-4),
    ("init_seed", 0),
    ("num_experiments", 1),
    ("num_epochs", 200),
    ("log_interval", 0.25),
    ("device", "cuda"),
    ("world_size", 1),
    ("resume_mode", 0),
    ("save_format", "pdf"),
])


def parse_control_name(control_name: str) -> "OrderedDict[str, str]":
    """Split a control string into its 9 named fields (all kept as strings, like the reference)."""
    parts = control_name.split("_")
    if len(parts) != len(CONTROL_KEYS):
        raise ValueError("control_name needs %d '_'-separated fields (%s), got %r"
                         % (len(CONTROL_KEYS), "_".join(CONTROL_KEYS), control_name))
    return OrderedDict(zip(CONTROL_KEYS, parts))


def format_control_name(control: Dict[str, str]) -> str:
    return "_".join(str(control[k]) for k in CONTROL_KEYS)


def parse_model_mode(model_mode: str):
    """``'a2-b8'`` -> ([1.0, 0.5], [2, 8], ['a', 'b'])."""
    rates, props, levels = [], [], []
    for tok in model_mode.split("-"):
        if not tok or tok[0] not in MODEL_SPLIT_RATE:
            raise ValueError("Not valid model mode token %r" % tok)
        levels.append(tok[0])
        rates.append(MODEL_SPLIT_RATE[tok[0]])
        props.append(int(tok[1:]))
    return rates, props, levels


def _fix_rates(rates: List[float], props: List[int], num_users: int) -> List[float]:
    # ``num_users // sum(prop)`` users per unit of proportion, tail padded with the last rate
    # (reference src/utils.py:134-144).
    per_unit = num_users // sum(props)
    out: List[float] = []
    for r, p in zip(rates, props):
        out += [r] * (per_unit * p)
    out += [out[-1]] * (num_users - len(out)) if out else [rates[-1]] * num_users
    return out[:num_users] if len(out) > num_users else out


def process_control(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """Derive every typed field the reference derives from ``cfg['control']`` (in place; also returned)."""
    c = cfg["control"]
    cfg["model_split_rate"] = dict(MODEL_SPLIT_RATE)
    cfg["fed"] = int(c["fed"])
    cfg["num_users"] = int(c["num_users"])
    cfg["frac"] = float(c["frac"])
    cfg["data_split_mode"] = c["data_split_mode"]
    cfg["model_split_mode"] = c["model_split_mode"]
    cfg["model_mode"] = c["model_mode"]
    cfg["norm"] = c["norm"]
    cfg["scale"] = bool(int(c["scale"]))
    cfg["mask"] = bool(int(c["mask"]))
    if cfg["norm"] not in ("bn", "in", "ln", "gn", "none"):
        raise ValueError("Not valid norm")
    cfg["global_model_mode"] = cfg["model_mode"][0]
    cfg["global_model_rate"] = MODEL_SPLIT_RATE[cfg["global_model_mode"]]
    rates, props, _ = parse_model_mode(cfg["model_mode"])
    if cfg["model_split_mode"] == "dynamic":
        cfg["model_rate"] = rates
        s = float(sum(props))
        cfg["proportion"] = [p / s for p in props]
    elif cfg["model_split_mode"] == "fix":
        cfg["model_rate"] = _fix_rates(rates, props, cfg["num_users"])
    else:
        raise ValueError("Not valid model split mode")
    cfg["conv"] = {"hidden_size": [64, 128, 256, 512]}
    cfg["resnet"] = {"hidden_size": [64, 128, 256, 512]}
    cfg["transformer"] = {"embedding_size": 256, "num_heads": 8, "hidden_size": 512, "num_layers": 4,
                          "dropout": 0.2}
    dn, dsm = cfg["data_name"], cfg["data_split_mode"]
    cfg["optimizer_name"] = "SGD"
    cfg["momentum"] = 0.9
    cfg["weight_decay"] = 5e-4
    cfg["scheduler_name"] = "MultiStepLR"
    cfg["factor"] = 0.1
    if dn in ("MNIST", "EMNIST", "FashionMNIST"):
        cfg["data_shape"] = [1, 28, 28]
        cfg["lr"] = 1e-2
        table = {"iid": ({"global": 200, "local": 5}, {"train": 10, "test": 50}, [100]),
                 "non-iid": ({"global": 400, "local": 5}, {"train": 10, "test": 50}, [200]),
                 "none": (200, {"train": 100, "test": 500}, [100])}
    elif dn in ("CIFAR10", "CIFAR100"):
        cfg["data_shape"] = [3, 32, 32]
        cfg["lr"] = 1e-1
        table = {"iid": ({"global": 400, "local": 5}, {"train": 10, "test": 50}, [150, 250]),
                 "non-iid": ({"global": 800, "local": 5}, {"train": 10, "test": 50}, [300, 500]),
                 "none": (400, {"train": 100, "test": 500}, [150, 250])}
    elif dn in LM_DATA:
        cfg["lr"] = 1e-1
        cfg["bptt"] = 64
        cfg["mask_rate"] = 0.15
        table = {"iid": ({"global": 200, "local": 1}, {"train": 100, "test": 10}, [50, 100]),
                 "none": (100, {"train": 100, "test": 100}, [25, 50])}
    else:
        raise ValueError("Not valid dataset")
    key = "non-iid" if "non-iid" in dsm else dsm
    if key not in table:
        raise ValueError("Not valid data_split_mode")
    ne, bs, ms = table[key]
    cfg["num_epochs"] = copy.deepcopy(ne)
    cfg["batch_size"] = dict(bs)
    cfg["milestones"] = list(ms)
    if "non-iid" in dsm:
        cfg["non-iid-n"] = int(dsm.split("-")[-1])
    return cfg


def make_cfg(data_name: str = "CIFAR10", model_name: str = "resnet18", control_name: str | None = None,
             **overrides: Any) -> Dict[str, Any]:
    """Build a fully processed config dict (defaults -> overrides -> control_name -> hyper-parameter table)."""
    cfg = copy.deepcopy(DEFAULT_CFG)
    cfg["data_name"] = data_name
    cfg["model_name"] = model_name
    post = {}
    for k, v in overrides.items():
        if k in DEFAULT_CFG:
            cfg[k] = v
        else:
            post[k] = v  # keys applied after the table (engine knobs, shortened schedules for tests/bench)
    if control_name:
        cfg["control"] = parse_control_name(control_name)
    cfg["control_name"] = format_control_name(cfg["control"])
    process_control(cfg)
    for k, v in post.items():
        cfg[k] = v
    is_lm = data_name in LM_DATA
    if cfg["fed"]:
        if is_lm:
            cfg["pivot_metric"], cfg["pivot"] = "Global-Perplexity", float("inf")
            cfg["metric_name"] = {"train": {"Local": ["Local-Loss", "Local-Perplexity"]},
                                  "test": {"Global": ["Global-Loss", "Global-Perplexity"]}}
        else:
            cfg["pivot_metric"], cfg["pivot"] = "Global-Accuracy", -float("inf")
            cfg["metric_name"] = {"train": {"Local": ["Local-Loss", "Local-Accuracy"]},
                                  "test": {"Local": ["Local-Loss", "Local-Accuracy"],
                                           "Global": ["Global-Loss", "Global-Accuracy"]}}
    else:
        if is_lm:
            cfg["pivot_metric"], cfg["pivot"] = "Perplexity", float("inf")
            cfg["metric_name"] = {"train": ["Loss", "Perplexity"], "test": ["Loss", "Perplexity"]}
        else:
            cfg["pivot_metric"], cfg["pivot"] = "Accuracy", -float("inf")
            cfg["metric_name"] = {"train": ["Loss", "Accuracy"], "test": ["Loss", "Accuracy"]}
    return cfg


def model_tag(cfg: Dict[str, Any], seed: int) -> str:
    """``{seed}_{data}_{subset}_{model}_{control_name}`` (reference src/train_classifier_fed.py:41-42)."""
    parts = [str(seed), cfg["data_name"], cfg["subset"], cfg["model_name"], cfg["control_name"]]
    return "_".join(x for x in parts if x)


def num_active_users(cfg: Dict[str, Any]) -> int:
    return int(math.ceil(cfg["frac"] * cfg["num_users"]))


def level_of_rate(rate: float) -> str:
    for k, v in MODEL_SPLIT_RATE.items():
        if abs(v - rate) < 1e-12:
            return k
    raise ValueError("rate %r is not a complexity level" % rate)
