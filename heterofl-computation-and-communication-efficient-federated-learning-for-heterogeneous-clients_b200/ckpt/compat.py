"""Reference-compatible (de)serialisation.

Checkpoints are ``torch.save(obj, path, pickle_protocol=2)`` dicts (reference ``src/utils.py:27-46``) that pickle a
``logger.Logger`` instance and -- for language models -- a ``datasets.lm.Vocab`` inside ``cfg``.  Those module paths
only exist when the reference's ``src/`` is the working directory, so :func:`compat_modules` installs tiny alias
modules under those names *for the duration of a save/load* (and removes them again): files written here unpickle in
the reference, and files written by the reference unpickle here.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch


@contextlib.contextmanager
def compat_modules():
    from ..data.datasets import Vocab
    from ..utils.logger import Logger
    Vocab.__module__ = "datasets.lm"
    installed = []
    swapped = []                       # (module, attr, previous value) -- e.g. the reference's own ``logger`` is imported

    def _install(name, **attrs):
        if name in sys.modules:
            mod = sys.modules[name]
            for k, v in attrs.items():
                swapped.append((mod, k, mod.__dict__.get(k, swapped)))
                setattr(mod, k, v)
            return
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        installed.append(name)

    _install("logger", Logger=Logger)
    _install("datasets")
    _install("datasets.lm", Vocab=Vocab)
    if "datasets" in installed:
        sys.modules["datasets"].lm = sys.modules["datasets.lm"]
    try:
        yield
    finally:
        for name in installed:
            sys.modules.pop(name, None)
        for mod, k, old in reversed(swapped):
            if old is swapped:
                delattr(mod, k)
            else:
                setattr(mod, k, old)


def save(obj, path, protocol: int = 2, mode: str = "torch") -> None:
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    # write-then-rename: a kill during the write must never truncate the only resume point (cli/supervise.py restarts
    # from this file); os.replace is atomic on POSIX
    tmp = path + ".tmp"
    if mode == "torch":
        with compat_modules():
            torch.save(obj, tmp, pickle_protocol=protocol)
        os.replace(tmp, path)
    elif mode == "numpy":
        with open(tmp, "wb") as f:
            np.save(f, obj, allow_pickle=True)
        os.replace(tmp, path)
    else:
        raise ValueError("Not valid save mode")


def atomic_copy(src: str, dst: str) -> None:
    """Copy ``src`` over ``dst`` without ever exposing a partially written ``dst``."""
    import shutil
    tmp = dst + ".tmp"
    shutil.copy(src, tmp)
    os.replace(tmp, dst)


def load(path, mode: str = "torch"):
    if mode == "torch":
        with compat_modules():
            return torch.load(path, map_location="cpu", weights_only=False)
    if mode == "numpy":
        return np.load(path, allow_pickle=True)
    raise ValueError("Not valid save mode")
