"""Importable alias for the framework package.

The product directory carries the (un-importable, hyphenated) project name required by the
build brief; ``hfl_b200`` simply points its ``__path__`` there so ``import hfl_b200.fed``
resolves to ``<project-dir>/fed``.
"""
import os as _os

PROJECT_DIR_NAME = ("heterofl-computation-and-communication-efficient-federated-learning-"
                    "for-heterogeneous-clients_b200")
_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PROJECT_DIR = _os.path.join(_root, PROJECT_DIR_NAME)
REPO_ROOT = _root
__path__.append(PROJECT_DIR)
__version__ = "0.1.0"
