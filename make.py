#!/usr/bin/env python
"""`python make.py --run train --model resnet18 --fed 1 --num_gpu 8 --data_split_mode non-iid-2` (reference src/make.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hfl_b200.utils.make import main  # noqa: E402

if __name__ == "__main__":
    main()
