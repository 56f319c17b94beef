#!/usr/bin/env python
"""`python train_transformer_fed.py --data_name ... --model_name ... --control_name ...` -- same CLI as the reference's src/train_transformer_fed.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hfl_b200.cli.main import run  # noqa: E402

if __name__ == "__main__":
    run("train_transformer_fed")
