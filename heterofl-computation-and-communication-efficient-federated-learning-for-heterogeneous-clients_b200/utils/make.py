"""Experiment-grid generators (reference ``src/make.py`` / ``src/make_ablation.py``).

Builds the control-name grids of the paper -- the 5 single levels, all 26 multi-level ``dynamic`` combinations, the 90
two-level ``fix`` interpolations ``x{i}-y{10-i}``, and the norm x scale x mask ablation grid -- and writes a bash script
that runs one experiment per GPU group (``CUDA_VISIBLE_DEVICES`` round-robin, ``wait`` every ``--round`` jobs).
``launcher='torchrun'`` instead emits one multi-GPU job per line (clients of a round sharded over ``world_size`` GPUs).
"""
from __future__ import annotations

import itertools
from typing import List

LEVELS = ["a", "b", "c", "d", "e"]
MODEL_DATA = {"conv": ("MNIST", "classifier"), "resnet18": ("CIFAR10", "classifier"),
              "transformer": ("WikiText2", "transformer")}


def combinations(levels: List[str]) -> List[str]:
    """All multi-level subsets as ``a1-b1-...`` (single levels dropped), in itertools order."""
    toks = [x + "1" for x in levels]
    out = []
    for i in range(1, len(toks) + 1):
        out.extend("-".join(x) for x in itertools.combinations(toks, i))
    return out[len(levels):]


def interpolations(levels: List[str]) -> List[str]:
    out = []
    for i in range(1, 10):
        for j in range(len(levels)):
            for k in range(j + 1, len(levels)):
                out.append("{}{}-{}{}".format(levels[j], i, levels[k], 10 - i))
    return out


def control_grid(fed: int, data_split_mode: str) -> List[str]:
    if fed == 0:
        groups = [[["0"], ["1"], ["1"], ["none"], ["fix"], [l + "1" for l in LEVELS], ["bn"], ["1"], ["1"]]]
    elif fed == 1:
        head = [["1"], ["100"], ["0.1"], [data_split_mode]]
        tail = [["bn"], ["1"], ["1"]]
        groups = [head + [["fix"], [l + "1" for l in LEVELS]] + tail,
                  head + [["dynamic"], combinations(LEVELS)] + tail,
                  head + [["fix"], interpolations(LEVELS)] + tail]
    else:
        raise ValueError("Not valid fed")
    out = []
    for g in groups:
        out.extend("_".join(x) for x in itertools.product(*g))
    return out


def ablation_grid(data_split_mode: str) -> List[str]:
    levels = ["a", "e"]
    comb = combinations(levels)
    norm_1, norm_2 = ["bn", "none"], ["in", "ln", "gn"]
    head = [["1"], ["100"], ["0.1"], [data_split_mode]]
    if data_split_mode == "iid":
        groups = [head + [["fix"], ["a1", "e1"], norm_2 + norm_1, ["1"], ["1"]],
                  head + [["dynamic"], comb, norm_2, ["1"], ["1"]],
                  head + [["dynamic"], comb, norm_1, ["0", "1"], ["1"]]]
    elif data_split_mode == "non-iid-2":
        groups = [head + [["fix"], ["a1", "e1"], norm_2, ["1"], ["1"]],
                  head + [["fix"], ["a1", "e1"], norm_1, ["1"], ["0", "1"]],
                  head + [["dynamic"], comb, norm_2, ["1"], ["1"]],
                  head + [["dynamic"], comb, norm_1, ["0", "1"], ["0", "1"]]]
    else:
        raise ValueError("Not valid data_split_mode for the ablation grid")
    out = []
    for g in groups:
        out.extend("_".join(x) for x in itertools.product(*g))
    return out


def make_script(run: str, model: str, fed: int, num_gpu: int, world_size: int, init_seed: int, round_: int,
                experiment_step: int, num_experiments: int, resume_mode: int, data_split_mode: str,
                ablation: bool = False, launcher: str = "python") -> (str, str):
    data_split_mode = data_split_mode if fed != 0 else "none"
    if model not in MODEL_DATA:
        raise ValueError("Not valid model")
    data_name, file = MODEL_DATA[model]
    script = "{}_{}{}.py".format(run, file, {0: "", 1: "_fed", 2: "_local"}[fed])
    gpu_ids = [",".join(str(i) for i in range(x, x + world_size)) for x in range(0, num_gpu, world_size)]
    seeds = list(range(init_seed, init_seed + num_experiments, experiment_step))
    controls = ablation_grid(data_split_mode) if ablation else control_grid(fed, data_split_mode)
    s = "#!/bin/bash\n"
    k = 0
    for seed in seeds:
        for c in controls:
            if launcher == "torchrun":
                cmd = ("CUDA_VISIBLE_DEVICES=\"{}\" python -m torch.distributed.run --nnodes=1 --nproc-per-node {} "
                       "--master-addr 127.0.0.1 --master-port {} {}").format(gpu_ids[k % len(gpu_ids)], world_size,
                                                                              29500 + k % 400, script)
            else:
                cmd = "CUDA_VISIBLE_DEVICES=\"{}\" python {}".format(gpu_ids[k % len(gpu_ids)], script)
            s += ("{} --data_name {} --model_name {} --init_seed {} --world_size {} --num_experiments {} "
                  "--resume_mode {} --control_name {}&\n").format(cmd, data_name, model, seed, world_size,
                                                                  experiment_step, resume_mode, c)
            if k % round_ == round_ - 1:
                s = s[:-2] + "\nwait\n"
            k += 1
    if s[-5:-1] != "wait":
        s += "wait\n"
    return "{}_{}_{}".format(run, model, data_split_mode), s


def main(argv=None, ablation: bool = False) -> None:
    import argparse
    ap = argparse.ArgumentParser(description="Config")
    ap.add_argument("--run", default="train", type=str)
    ap.add_argument("--model", default=None, type=str)
    ap.add_argument("--fed", default=1, type=int)
    ap.add_argument("--num_gpu", default=4, type=int)
    ap.add_argument("--world_size", default=1, type=int)
    ap.add_argument("--init_seed", default=0, type=int)
    ap.add_argument("--round", default=4, type=int)
    ap.add_argument("--experiment_step", default=1, type=int)
    ap.add_argument("--num_experiments", default=1, type=int)
    ap.add_argument("--resume_mode", default=0, type=int)
    ap.add_argument("--data_split_mode", default="iid", type=str)
    ap.add_argument("--launcher", default="python", choices=["python", "torchrun"])
    a = vars(ap.parse_args(argv))
    name, s = make_script(a["run"], a["model"], a["fed"], a["num_gpu"], a["world_size"], a["init_seed"], a["round"],
                          a["experiment_step"], a["num_experiments"], a["resume_mode"], a["data_split_mode"], ablation,
                          a["launcher"])
    print(s)
    with open("./{}.sh".format(name), "w") as f:
        f.write(s)
