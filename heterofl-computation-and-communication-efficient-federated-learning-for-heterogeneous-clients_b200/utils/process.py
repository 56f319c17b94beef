"""Result post-processing (reference ``src/process.py``).

Walks the experiment grids of :mod:`hfl_b200.utils.make`, loads ``output/result/{tag}.pt`` files written by the
``test_*`` scripts (``{cfg, epoch, logger: {train, test}}``), nests the final test metrics and the per-round histories by
control field, reduces over seeds (mean / std / min / max), mixes the per-level parameter / FLOP / size figures from
``summary.py`` by the control's level proportions (``make_stats``), and exports

* ``processed_result_exp.json`` / ``.pt`` and ``processed_result_history.pt``,
* a result table per (data, split) -- ``result.xlsx`` when an Excel engine is installed, ``result.csv`` otherwise,
* interpolation plots (metric vs. number of parameters) and learning curves when matplotlib is available.
"""
from __future__ import annotations

import itertools
import json
import os
from collections import defaultdict
from typing import Dict, List

import numpy as np

from ..ckpt.compat import load, save
from .make import LEVELS, combinations, interpolations

DATA_MODEL = {"MNIST": "conv", "CIFAR10": "resnet18", "WikiText2": "transformer"}
DATA_SPLITS = {"MNIST": ["iid", "non-iid-2"], "CIFAR10": ["iid", "non-iid-2"], "WikiText2": ["iid"]}
METRIC_NAME = {"MNIST": "Accuracy", "CIFAR10": "Accuracy", "WikiText2": "Perplexity"}
INTERP_NAME = ["a-b", "a-c", "a-d", "a-e", "b-c", "b-d", "b-e", "c-d", "c-e", "d-e", "e"]
MARKERS = dict(zip(INTERP_NAME, ["o", "s", "p", "P", "*", "h", "X", "D", "^", "v", ">"]))


def make_control_list(data_name: str, exp: List[str]) -> List[List[List[str]]]:
    model = DATA_MODEL[data_name]
    singles = [l + "1" for l in LEVELS]
    head = [exp, [data_name], ["label"], [model]]
    fed = ["1"], ["100"], ["0.1"], DATA_SPLITS[data_name]
    tail = [["bn"], ["1"], ["1"]]
    return [head + [["0"], ["1"], ["1"], ["none"], ["fix"], singles] + tail,
            head + list(fed) + [["fix"], singles] + tail,
            head + list(fed) + [["dynamic"], combinations(LEVELS)] + tail,
            head + list(fed) + [["fix"], interpolations(LEVELS)] + tail]


def make_ablation_control_list(data_name: str, exp: List[str]) -> List[List[List[str]]]:
    model = DATA_MODEL[data_name]
    n1, n2, comb = ["bn", "none"], ["in", "ln", "gn"], ["a1-e1"]
    out = []
    for d in [s for s in DATA_SPLITS[data_name] if data_name != "WikiText2"]:
        h = [exp, [data_name], ["label"], [model], ["1"], ["100"], ["0.1"], [d]]
        if d == "iid":
            out += [h + [["fix"], ["a1", "e1"], n2 + n1, ["1"], ["1"]], h + [["dynamic"], comb, n2, ["1"], ["1"]],
                    h + [["dynamic"], comb, n1, ["0", "1"], ["1"]]]
        else:
            out += [h + [["fix"], ["a1", "e1"], n2, ["1"], ["1"]], h + [["fix"], ["a1", "e1"], n1, ["1"], ["0", "1"]],
                    h + [["dynamic"], comb, n2, ["1"], ["1"]], h + [["dynamic"], comb, n1, ["0", "1"], ["0", "1"]]]
    return out


def make_stats(model_tag: str, result_path: str):
    """(ratio, params, flops, space) of a control = proportion-weighted mix of its levels' summary results."""
    parts = model_tag.split("_")
    data_name, model_name, modes = parts[1], parts[3], parts[-4].split("-")
    g = load(os.path.join(result_path, "{}_{}_{}.pt".format(data_name, model_name, modes[0][0])))["num_params"]
    tot = dict(p=0.0, f=0.0, s=0.0, g=0.0, w=0)
    for m in modes:
        w = int(m[1:])
        st = load(os.path.join(result_path, "{}_{}_{}.pt".format(data_name, model_name, m[0])))
        tot["p"] += st["num_params"] * w
        tot["f"] += st["num_flops"] * w
        tot["s"] += st["space"] * w
        tot["g"] += g * w
        tot["w"] += w
    return tot["p"] / tot["g"], tot["p"] / tot["w"], tot["f"] / tot["w"], tot["s"] / tot["w"]


def _insert(tree_exp, tree_hist, control: List[str], tag: str, exp: List[str], result_path: str, missing: List[str]):
    node_e, node_h = tree_exp, tree_hist
    for key in control[1:]:
        node_e, node_h = node_e.setdefault(key, {}), node_h.setdefault(key, {})
    path = os.path.join(result_path, "%s.pt" % tag)
    if not os.path.exists(path):
        missing.append(path)
        return
    n = len(exp)
    i = exp.index(control[0])
    try:
        ratio, params, flops, space = make_stats(tag, result_path)
        for k, v in (("Ratio", ratio), ("Params", params), ("FLOPs", flops), ("Space", space)):
            node_e[k] = {"exp": [v]}
    except FileNotFoundError:
        pass
    res = load(path)
    for k in res["logger"]["test"].mean:
        name = k.split("/")[1]
        node_e.setdefault(name, {"exp": [None] * n})["exp"][i] = res["logger"]["test"].mean[k]
        node_h.setdefault(name, {"history": [None] * n})["history"][i] = res["logger"]["train"].history[k]


def summarize_result(node: Dict) -> None:
    for pivot in ("exp", "history"):
        if pivot in node:
            arr = np.stack([np.asarray(x, dtype=np.float64) for x in node[pivot] if x is not None], axis=0)
            scal = pivot == "exp"
            for name, fn in (("mean", np.mean), ("std", np.std), ("max", np.max), ("min", np.min),
                             ("argmax", np.argmax), ("argmin", np.argmin)):
                val = fn(arr, axis=0)
                node[name] = val.item() if scal else val
            node[pivot] = arr.tolist()
            return
    for v in node.values():
        if isinstance(v, dict):
            summarize_result(v)


def process_result(controls, exp: List[str], result_path: str):
    tree_e, tree_h, missing = {}, {}, []
    for c in controls:
        _insert(tree_e, tree_h, list(c), "_".join(c), exp, result_path, missing)
    summarize_result(tree_e)
    summarize_result(tree_h)
    return tree_e, tree_h, missing


def extract_processed_result(out: Dict, node: Dict, control: List[str]) -> None:
    if "exp" in node or "history" in node:
        name = "_".join(control[:-1])
        out.setdefault(name, {})
        out[name]["%s_mean" % control[-1]] = node.get("mean")
        out[name]["%s_std" % control[-1]] = node.get("std")
        return
    for k, v in node.items():
        if isinstance(v, dict):
            extract_processed_result(out, v, control + [k])


def make_df(extracted: Dict, result_path: str):
    import pandas as pd
    df = defaultdict(list)
    for exp_name, rec in extracted.items():
        control = exp_name.split("_")
        split, msplit, cname = control[-6:-3]
        rest = control[:-4] + control[-3:]
        frame = pd.DataFrame(data={k: [v] for k, v in rec.items()}, index=[cname])
        if split == "none" or not (msplit == "fix" or cname == "e1"):
            df["_".join(rest)].append(frame)
        elif "-" in cname:
            df["{}_{}".format("_".join(rest), "-".join(x[0] for x in cname.split("-")))].append(frame)
        else:
            for label in INTERP_NAME:
                if cname[0] == label[0]:
                    df["{}_{}".format("_".join(rest), label)].append(frame)
    out = {}
    for name, frames in df.items():
        t = pd.concat(frames)
        out[name] = t.sort_values(by=["Params_mean"], ascending=False) if "Params_mean" in t else t
    try:
        with pd.ExcelWriter(os.path.join(result_path, "result.xlsx")) as writer:
            row = 0
            for name, t in out.items():
                t.to_excel(writer, sheet_name="Sheet1", startrow=row + 1)
                writer.sheets["Sheet1"].cell(row=row + 1, column=1, value=name) if hasattr(
                    writer.sheets["Sheet1"], "cell") else writer.sheets["Sheet1"].write_string(row, 0, name)
                row += len(t.index) + 3
    except Exception:
        with open(os.path.join(result_path, "result.csv"), "w") as f:
            for name, t in out.items():
                f.write(name + "\n" + t.to_csv() + "\n")
    return out


def _plt():
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


def make_vis(df: Dict, vis_path: str, save_format: str = "pdf") -> int:
    plt = _plt()
    if plt is None:
        print("matplotlib is not installed: skipping interpolation plots")
        return 0
    figs = {}
    for name, t in df.items():
        if "fix" not in name or "none" in name or "Params_mean" not in t:
            continue
        control = name.split("_")
        metric, label = METRIC_NAME[control[0]], control[-1]
        x = t["Params_mean"]
        kinds = [("local", "Local-"), ("global", "Global-")] if "non-iid-2" in name else [("", "Global-")]
        for suffix, pre in kinds:
            col = "%s%s_mean" % (pre, metric)
            if col not in t:
                continue
            fig_name = "{}_{}{}".format("_".join(control[:-1]), label[0], ("_" + suffix) if suffix else "")
            figs[fig_name] = plt.figure(fig_name)
            plt.plot(x, t[col], linestyle="-", marker=MARKERS.get(label, "o"), label=label)
            plt.legend(fontsize=16)
            plt.xlabel("Number of Model Parameters", fontsize=16)
            plt.ylabel(metric, fontsize=16)
            plt.ticklabel_format(axis="x", style="sci", scilimits=(0, 0))
    os.makedirs(vis_path, exist_ok=True)
    for fig_name in figs:
        plt.figure(fig_name)
        plt.grid()
        plt.savefig(os.path.join(vis_path, "%s.%s" % (fig_name, save_format)), dpi=500, bbox_inches="tight", pad_inches=0)
        plt.close(fig_name)
    return len(figs)


def make_learning_curve(extracted_hist: Dict, vis_path: str, save_format: str = "pdf") -> int:
    plt = _plt()
    if plt is None:
        print("matplotlib is not installed: skipping learning curves")
        return 0
    figs = {}
    wanted = {"a5-b5", "a5-c5", "a5-d5", "a5-e5", "a1-b1", "a1-c1", "a1-d1", "a1-e1"}
    for exp_name, rec in extracted_hist.items():
        control = exp_name.split("_")
        cname = control[-4]
        if cname not in wanted:
            continue
        metric = METRIC_NAME[control[0]]
        label = "-".join(x[0] for x in cname.split("-"))
        kinds = [("local", "Local-"), ("global", "Global-")] if "non-iid-2" in exp_name else [("global", "Global-")]
        for suffix, pre in kinds:
            y = rec.get("%s%s_mean" % (pre, metric))
            if y is None:
                continue
            fig_name = "{}_lc_{}".format("_".join(control[:-4] + control[-3:]), suffix)
            figs[fig_name] = plt.figure(fig_name)
            plt.plot(np.arange(len(y)), y, linestyle="-", label=label)
            plt.legend(fontsize=16)
            plt.xlabel("Communication rounds", fontsize=16)
            plt.ylabel("Test {}".format(metric), fontsize=16)
    os.makedirs(vis_path, exist_ok=True)
    for fig_name in figs:
        plt.figure(fig_name)
        plt.grid()
        plt.savefig(os.path.join(vis_path, "%s.%s" % (fig_name, save_format)), dpi=500, bbox_inches="tight", pad_inches=0)
        plt.close(fig_name)
    return len(figs)


def main(argv=None) -> None:
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--result_path", default="./output/result")
    ap.add_argument("--vis_path", default="./output/vis")
    ap.add_argument("--num_experiments", default=1, type=int)
    ap.add_argument("--ablation", default=0, type=int)
    ap.add_argument("--data_names", default="MNIST,CIFAR10,WikiText2")
    ap.add_argument("--save_format", default="pdf")
    a = ap.parse_args(argv)
    exp = [str(x) for x in range(a.num_experiments)]
    lists = []
    for d in a.data_names.split(","):
        lists += make_ablation_control_list(d, exp) if a.ablation else make_control_list(d, exp)
    controls = []
    for l in lists:
        controls.extend(itertools.product(*l))
    tree_e, tree_h, missing = process_result(controls, exp, a.result_path)
    for m in missing:
        print("Missing {}".format(m))
    os.makedirs(a.result_path, exist_ok=True)
    with open(os.path.join(a.result_path, "processed_result_exp.json"), "w") as fp:
        json.dump(tree_e, fp, indent=2)
    save(tree_e, os.path.join(a.result_path, "processed_result_exp.pt"))
    save(tree_h, os.path.join(a.result_path, "processed_result_history.pt"))
    ext = {}
    extract_processed_result(ext, tree_e, [])
    df = make_df(ext, a.result_path) if ext else {}
    make_vis(df, a.vis_path, a.save_format)
    ext_h = {}
    extract_processed_result(ext_h, tree_h, [])
    make_learning_curve(ext_h, a.vis_path, a.save_format)
