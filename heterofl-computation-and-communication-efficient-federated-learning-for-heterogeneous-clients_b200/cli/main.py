"""Command-line entry points with the reference's interface.

Every key of the default config is a ``--flag`` (typed by its default, like the reference's ``exec``-generated parser,
``src/train_classifier_fed.py:20-30``) plus ``--control_name``.  Scripts (thin wrappers at the repo root):

    train_classifier_fed.py / train_transformer_fed.py    federated training           (reference C31 / C32)
    train_classifier.py / train_transformer.py            centralised baselines        (C33 / C34)
    test_classifier_fed.py / test_transformer_fed.py      evaluate the *_best checkpoint -> output/result/{tag}.pt (C35/C36)
    test_classifier.py / test_transformer.py              centralised evaluation       (C37)

Extra engine flags (not in the reference): ``--synthetic {auto,1,0}``, ``--out_dir``, ``--compute_dtype``,
``--max_concurrent``, ``--agg_backend``, ``--rounds`` (stop early).  Multi-GPU: launch with ``torchrun``.
"""
from __future__ import annotations

import argparse
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # many concurrent client streams (set before CUDA init)
from typing import Dict, List, Optional

import torch

from ..ckpt import compat as ckpt
from ..config.control import DEFAULT_CFG, LM_DATA, make_cfg, model_tag
from ..parallel.dist import init_distributed, shutdown
from ..utils.logger import Logger

SCRIPTS = ("train_classifier_fed", "train_transformer_fed", "train_classifier", "train_transformer",
           "test_classifier_fed", "test_transformer_fed", "test_classifier", "test_transformer")


def build_parser(script: str) -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="cfg", prog=script + ".py")
    for k, v in DEFAULT_CFG.items():
        if k == "control":
            continue
        if isinstance(v, (dict, list)):
            ap.add_argument("--" + k, default=None, type=str)      # structured keys: not overridable (as in the reference)
        elif isinstance(v, bool):
            ap.add_argument("--" + k, default=v, type=lambda s: str(s).lower() in ("1", "true", "yes"))
        else:
            ap.add_argument("--" + k, default=v, type=type(v))
    ap.add_argument("--control_name", default=None, type=str)
    # like the reference, a missing dataset is an error by default: synthetic data must be asked for (--synthetic 1;
    # "auto" = fall back with a loud warning, recorded as cfg["synthetic_data"])
    ap.add_argument("--synthetic", default="0", type=str)
    ap.add_argument("--out_dir", default="./output", type=str)
    ap.add_argument("--compute_dtype", default="auto", type=str)
    ap.add_argument("--max_concurrent", default=16, type=int)
    ap.add_argument("--agg_backend", default="auto", type=str)
    ap.add_argument("--rounds", default=0, type=int, help="run at most this many rounds/epochs (0 = full schedule)")
    ap.add_argument("--quiet", default=0, type=int)
    ap.add_argument("--nvtx", default=0, type=int, help="emit an NVTX range per round phase")
    ap.add_argument("--fault_drop", default=0.0, type=float,
                    help="fault injection: probability that an active client's update is lost each round")
    ap.add_argument("--async_eval", default="auto", type=str,
                    help="overlap a round's sBN + test passes with the next round's training: auto (vision models from 4 "
                         "ranks up) | 1 | 0; test metrics are then logged one round late")
    return ap


def cfg_from_args(script: str, args: Dict) -> Dict:
    is_lm = "transformer" in script
    fed = script.endswith("_fed")
    data_name = args["data_name"]
    model_name = args["model_name"]
    if is_lm and data_name not in LM_DATA:
        data_name = "WikiText2"
    if is_lm:
        model_name = "transformer"
    control = args["control_name"]
    if control is None:
        control = "1_100_0.1_iid_fix_a1_bn_1_1" if fed else "0_1_1_none_fix_a1_bn_1_1"
    over = {k: args[k] for k in ("subset", "num_workers", "init_seed", "num_experiments", "log_interval", "device",
                                 "world_size", "resume_mode", "save_format") if args.get(k) is not None}
    syn = args["synthetic"]
    over["synthetic"] = True if syn in ("1", "true", "True") else (False if syn in ("0", "false", "False") else "auto")
    return make_cfg(data_name, model_name, control, **over)


def _options(args: Dict):
    from ..fed.engine import EngineOptions
    return EngineOptions(compute_dtype=args["compute_dtype"], max_concurrent=args["max_concurrent"],
                         agg_backend=args["agg_backend"], out_dir=args["out_dir"], tensorboard=True,
                         quiet=bool(args["quiet"]), nvtx=bool(args.get("nvtx", 0)),
                         fault_drop=float(args.get("fault_drop", 0.0)),
                         async_eval={"1": True, "true": True, "0": False, "false": False}.get(
                             str(args.get("async_eval", "auto")).lower(), None))


def run(script: str, argv: Optional[List[str]] = None) -> None:
    assert script in SCRIPTS, script
    args = vars(build_parser(script).parse_args(argv))
    device = args["device"] if (args["device"] != "cuda" or torch.cuda.is_available()) else "cpu"
    info = init_distributed(device)
    base_cfg = cfg_from_args(script, args)
    seeds = list(range(base_cfg["init_seed"], base_cfg["init_seed"] + base_cfg["num_experiments"]))
    for seed in seeds:
        cfg = cfg_from_args(script, args)
        cfg["model_tag"] = model_tag(cfg, seed)
        if info.is_main:
            print("Experiment: {}".format(cfg["model_tag"]))
        if script.startswith("train_") and script.endswith("_fed"):
            from ..fed.engine import Engine
            eng = Engine(cfg, info, options=_options(args), seed=seed)
            eng.setup()
            eng.fit(args["rounds"] or None)
        elif script.startswith("train_"):
            from ..fed.central import CentralTrainer
            tr = CentralTrainer(cfg, info, options=_options(args), seed=seed)
            tr.fit(args["rounds"] or None)
        else:
            evaluate(script, cfg, info, args, seed)
    shutdown()


def evaluate(script: str, cfg: Dict, info, args: Dict, seed: int) -> Dict:
    """``test_*`` scripts: load ``{tag}_best.pt``, (vision) re-estimate sBN statistics with the TEST batch size (the
    reference sets ``batch_size['train'] = batch_size['test']`` first, ``src/test_classifier_fed.py:42``), evaluate,
    and write ``output/result/{tag}.pt = {cfg, epoch, logger: {train, test}}``."""
    from ..fed.central import CentralTrainer
    from ..fed.engine import Engine
    opts = _options(args)
    opts.save = False
    out_dir = opts.out_dir
    cfg["batch_size"]["train"] = cfg["batch_size"]["test"]
    best = os.path.join(out_dir, "model", "%s_best.pt" % cfg["model_tag"])
    last = os.path.join(out_dir, "model", "%s_checkpoint.pt" % cfg["model_tag"])
    if not os.path.exists(best):
        raise FileNotFoundError(best)
    ck = ckpt.load(best)
    fed = script.endswith("_fed")
    test_logger = Logger(os.path.join(out_dir, "runs", "test_%s" % cfg["model_tag"]))
    test_logger.safe(True)
    if fed:
        eng = Engine(cfg, info, options=opts, seed=seed)
        eng.data_split, eng.label_split = ck["data_split"], ck["label_split"]
        eng.load_state_dict(ck["model_dict"], strict=False)
        eng.setup(resume_mode=0)
        gen = torch.Generator().manual_seed(seed)
        if eng.is_lm:
            res = eng.evaluator.test_lm(eng.global_arena, eng.test_rows.to(eng.device))
        else:
            running = eng.evaluator.sbn_stats(eng.global_arena, eng.dataset["train"], eng.train_img.to(eng.device), gen)
            res = eng.evaluator.test_vision(eng.global_arena, running, eng.dataset["test"], eng.test_img.to(eng.device),
                                            eng.test_lab.to(eng.device), eng.test_owner, eng.user_masks)
        names = cfg["metric_name"]["test"].get("Local", []) + cfg["metric_name"]["test"]["Global"]
        loc = {k: v for k, v in res.items() if k.startswith("Local-")}
        glo = {k: v for k, v in res.items() if k.startswith("Global-")}
        if loc:
            test_logger.append(loc, "test", n=max(res.get("_n_local", 1.0), 1.0))
        test_logger.append(glo, "test", n=max(res.get("_n_global", 1.0), 1.0))
    else:
        tr = CentralTrainer(cfg, info, options=opts, seed=seed)
        tr.load_state_dict(ck["model_dict"], strict=False)
        gen = torch.Generator().manual_seed(seed)
        if tr.is_lm:
            r = tr.evaluator.test_lm(tr.global_arena, tr.test_rows)
            res = {"Loss": r["Global-Loss"], "Perplexity": r["Global-Perplexity"]}
        else:
            running = tr.evaluator.sbn_stats(tr.global_arena, tr.dataset["train"], tr.train_img, gen)
            r = tr.evaluator.test_vision(tr.global_arena, running, tr.dataset["test"], tr.test_img, tr.test_lab, None, None)
            res = {"Loss": r["Global-Loss"], "Accuracy": r["Global-Accuracy"]}
        names = cfg["metric_name"]["test"]
        test_logger.append(res, "test", n=max(r["_n_global"], 1.0))
    test_logger.append({"info": ["Model: {}".format(cfg["model_tag"]), "Test Epoch: {}(100%)".format(ck["epoch"])]},
                       "test", mean=False)
    test_logger.write("test", names, quiet=not info.is_main)
    test_logger.safe(False)
    train_logger = ckpt.load(last)["logger"] if os.path.exists(last) else ck["logger"]
    result = {"cfg": dict(cfg), "epoch": ck["epoch"], "logger": {"train": train_logger, "test": test_logger}}
    if info.is_main:
        ckpt.save(result, os.path.join(out_dir, "result", "%s.pt" % cfg["model_tag"]))
    return result


if __name__ == "__main__":
    run(sys.argv[1], sys.argv[2:])
