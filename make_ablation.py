#!/usr/bin/env python
"""Ablation grid over norm x scaler x masked-CE (reference src/make_ablation.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hfl_b200.utils.make import main  # noqa: E402

if __name__ == "__main__":
    main(ablation=True)
