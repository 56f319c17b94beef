#!/usr/bin/env python
"""`python summary.py --data_name CIFAR10 --model_name resnet18 --control_name 1_100_0.1_iid_fix_a1_bn_1_1` -- parameter /
FLOP table of the model at the control's first level (reference src/summary.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    from hfl_b200.cli.main import build_parser, cfg_from_args
    from hfl_b200.data.datasets import DATA_SIZES, LM_SIZES
    from hfl_b200.utils.summary import run
    args = vars(build_parser("summary").parse_args())
    script = "train_transformer_fed" if args["model_name"] == "transformer" else "train_classifier_fed"
    cfg = cfg_from_args(script, args)
    if cfg["data_name"] in DATA_SIZES:
        cfg["classes_size"] = DATA_SIZES[cfg["data_name"]][2]
    else:
        cfg["num_tokens"] = LM_SIZES[cfg["data_name"]][2]
    run(cfg, args["out_dir"])
