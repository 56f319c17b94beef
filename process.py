#!/usr/bin/env python
"""`python process.py [--num_experiments N] [--ablation 1]` -- aggregate output/result/*.pt (reference src/process.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hfl_b200.utils.process import main  # noqa: E402

if __name__ == "__main__":
    main()
