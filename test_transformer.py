#!/usr/bin/env python
"""`python test_transformer.py --data_name ... --model_name ... --control_name ...` -- same CLI as the reference's src/test_transformer.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hfl_b200.cli.main import run  # noqa: E402

if __name__ == "__main__":
    run("test_transformer")
