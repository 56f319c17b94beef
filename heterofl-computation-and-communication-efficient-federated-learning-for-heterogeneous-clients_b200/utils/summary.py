"""Model summary: per-layer shapes, parameter counts and FLOPs of the sub-network at a complexity level.

Reference ``src/summary.py`` registers forward hooks on every ``nn.Module`` and counts, for batch size 1, parameters
(by mask) and FLOPs (``compute_flops`` :200-276: Conv2d MACs + bias, Norm 2*numel (BatchNorm/LayerNorm; GroupNorm is
not recognised -> 0), Linear MACs), then saves ``{num_params, num_flops, space}`` to
``output/result/{data}_{model}_{level}.pt`` for the params/FLOPs columns of the result tables.  Quirk kept for
parity: a module's FLOPs are added once **per parameter tensor** it owns (:146-148), so layers with weight+bias count
twice.  Here the same table is derived analytically from the :class:`ModelSpec` (no forward pass needed).
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

from ..plan.spec import ModelSpec, build_spec


def divide_by_unit(value: float) -> str:
    if value > 1e9:
        return "{:.6} G".format(value / 1e9)
    if value > 1e6:
        return "{:.6} M".format(value / 1e6)
    if value > 1e3:
        return "{:.6} K".format(value / 1e3)
    return "{:.6}".format(value / 1.0)


def _norm_flops(norm: str, numel: int) -> int:
    if norm == "bn":
        return 2 * numel
    return 0      # GroupNorm-based variants: "[Flops]: GroupNorm is not supported!" -> 0 in the reference


def layer_records(spec: ModelSpec, cfg: Dict, rate: float) -> List[Dict]:
    """One record per parameterised module: name, input/weight/output size, params, flops (batch 1)."""
    m = spec.meta
    recs: List[Dict] = []
    cnt: Dict[str, int] = {}

    def add(kind, inp, wsize, out, nparams, flops, ntensors):
        cnt[kind] = cnt.get(kind, 0) + 1
        recs.append(dict(module_name="%s_%d" % (kind, cnt[kind]), input_size=[list(inp)], weight_size=list(wsize),
                         output_size=[list(out)], params=int(nparams), flops=int(flops), n_param_tensors=ntensors))

    if spec.family in ("conv", "resnet"):
        norm = m["norm"]
        norm_kind = "BatchNorm2d" if norm == "bn" else "GroupNorm"
        C0, H, W = m["data_shape"]
        hidden = [int(np.ceil(rate * h)) for h in m["hidden"]]
        classes = m["classes"]

        def conv(cin, cout, k, stride, pad, h, w, bias):
            oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
            fl = k * k * cin * cout * oh * ow + (cout * oh * ow if bias else 0)
            add("Conv2d", (1, cin, h, w), (cout, cin, k, k), (1, cout, oh, ow), cout * cin * k * k + (cout if bias else 0),
                fl, 2 if bias else 1)
            return oh, ow

        def normrec(c, h, w):
            if norm != "none":
                add(norm_kind, (1, c, h, w), (c,), (1, c, h, w), 2 * c, _norm_flops(norm, c * h * w), 2)

        if spec.family == "conv":
            cin, h, w = C0, H, W
            for i, c in enumerate(hidden):
                h, w = conv(cin, c, 3, 1, 1, h, w, True)
                normrec(c, h, w)
                if i < len(hidden) - 1:
                    h, w = h // 2, w // 2
                cin = c
            add("Linear", (1, cin), (classes, cin), (1, classes), classes * cin + classes, cin * classes, 2)
        else:
            exp = m["expansion"]
            h, w = conv(C0, hidden[0], 3, 1, 1, H, W, False)
            cin = hidden[0]
            # module registration order inside a block: n1, conv1, n2, conv2, [n3, conv3], shortcut; hooks fire in
            # execution order n1, shortcut, conv1, n2, conv2 -- the table follows execution order like the reference
            for blk in m["blocks"]:
                planes = int(np.ceil(rate * blk["planes"]))
                stride = blk["stride"]
                normrec(cin, h, w)
                if blk["shortcut"]:
                    conv(cin, planes * exp, 1, stride, 0, h, w, False)
                if blk["kind"] == "block":
                    h2, w2 = conv(cin, planes, 3, stride, 1, h, w, False)
                    normrec(planes, h2, w2)
                    conv(planes, planes, 3, 1, 1, h2, w2, False)
                else:
                    conv(cin, planes, 1, 1, 0, h, w, False)
                    normrec(planes, h, w)
                    h2, w2 = conv(planes, planes, 3, stride, 1, h, w, False)
                    normrec(planes, h2, w2)
                    conv(planes, planes * exp, 1, 1, 0, h2, w2, False)
                h, w, cin = h2, w2, planes * exp
            normrec(cin, h, w)
            add("Linear", (1, cin), (classes, cin), (1, classes), classes * cin + classes, cin * classes, 2)
    else:
        V, S, H_ = m["num_tokens"], m["bptt"], m["heads"]
        E = int(np.ceil(rate * m["emb"]))
        Hd = int(np.ceil(rate * m["hidden"]))
        rows = 1

        def lin(cin, cout, name="Linear"):
            add(name, (rows, S, cin), (cout, cin), (rows, S, cout), cout * cin + cout, rows * S * cin * cout, 2)

        def ln(c):
            add("LayerNorm", (rows, S, c), (c,), (rows, S, c), 2 * c, 2 * rows * S * c, 2)

        add("Embedding", (rows, S), (S, E), (rows, S, E), S * E, 0, 1)
        add("Embedding", (rows, S), (V + 1, E), (rows, S, E), (V + 1) * E, 0, 1)
        ln(E)
        for _ in range(m["layers"]):
            for _q in range(4):
                lin(E, E)
            ln(E)
            lin(E, Hd)
            lin(Hd, E)
            ln(E)
        lin(E, E)
        ln(E)
        lin(E, V)
    return recs


def summarize(cfg: Dict, rate: float = None, spec: ModelSpec = None) -> Dict:
    spec = spec if spec is not None else build_spec(cfg)
    rate = cfg["global_model_rate"] if rate is None else rate
    recs = layer_records(spec, cfg, rate)
    num_params = sum(r["params"] for r in recs)
    num_flops = sum(r["flops"] * r["n_param_tensors"] for r in recs)      # reference double-count quirk
    macs = sum(r["flops"] for r in recs if r["module_name"].startswith(("Conv2d", "Linear")))
    return {"records": recs, "num_params": num_params, "num_flops": num_flops, "num_macs_exact": macs,
            "space": num_params * 32.0 / 8 / (1024 ** 2.0)}


def format_summary(summary: Dict) -> str:
    from tabulate import tabulate
    headers = ["Module Name", "Input Size", "Weight Size", "Output Size", "Parameters", "FLOPs"]
    rows = [[r["module_name"], str(r["input_size"]), str(r["weight_size"]), str(r["output_size"]), r["params"],
             divide_by_unit(r["flops"])] for r in summary["records"]]
    out = tabulate(rows, headers=headers, tablefmt="github") + "\n"
    out += "================================================================\n"
    out += "Total Number of Parameters: {} ({})\n".format(summary["num_params"], divide_by_unit(summary["num_params"]))
    out += "Total Number of FLOPs: {} ({})\n".format(summary["num_flops"], divide_by_unit(summary["num_flops"]))
    out += "Total Space (MB): {:.2f}\n".format(summary["space"])
    return out


def run(cfg: Dict, out_dir: str = "./output") -> Dict:
    """``summary.py`` entry point: print the table, write ``output/summary.md`` and the result ``.pt``."""
    from ..ckpt.compat import save
    s = summarize(cfg)
    text = format_summary(s)
    print(text)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "summary.md"), "w") as f:
        f.write(text)
    total = {"num_params": s["num_params"], "num_flops": s["num_flops"], "space": s["space"]}
    tag = "{}_{}_{}".format(cfg["data_name"], cfg["model_name"], cfg["model_mode"][0])
    save(total, os.path.join(out_dir, "result", "%s.pt" % tag))
    return total
