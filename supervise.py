"""Abort-and-resume wrapper: ``python supervise.py -- python train_classifier_fed.py ...`` (see hfl_b200/cli/supervise.py)."""
import sys

from hfl_b200.cli.supervise import main

if __name__ == "__main__":
    sys.exit(main())
