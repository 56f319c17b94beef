"""Nested-container helpers with the reference's names (``src/utils.py:12-24,56-97,347-350``): kept for user code that
imports them; the engine itself works on flat arenas and tensor datasets and does not need them."""
from __future__ import annotations

import collections.abc as container_abcs
import errno
import os
from itertools import repeat

import numpy as np
import torch


def check_exists(path: str) -> bool:
    return os.path.exists(path)


def makedir_exist_ok(path: str) -> None:
    try:
        os.makedirs(path)
    except OSError as e:
        if e.errno != errno.EEXIST:
            raise


def recur(fn, inp, *args):
    """Apply ``fn`` to every tensor / ndarray leaf of nested lists / tuples / dicts."""
    if isinstance(inp, (torch.Tensor, np.ndarray)):
        return fn(inp, *args)
    if isinstance(inp, list):
        return [recur(fn, x, *args) for x in inp]
    if isinstance(inp, tuple):
        return tuple(recur(fn, x, *args) for x in inp)
    if isinstance(inp, dict):
        return {k: recur(fn, v, *args) for k, v in inp.items()}
    raise ValueError("Not valid input type")


def to_device(inp, device):
    return recur(lambda x, d: x.to(d), inp, device)


def collate(inp: dict) -> dict:
    for k in inp:
        inp[k] = torch.stack(inp[k], 0)
    return inp


def ntuple(n: int):
    def parse(x):
        if isinstance(x, container_abcs.Iterable) and not isinstance(x, str):
            return x
        return tuple(repeat(x, n))
    return parse


def apply_fn(module, fn: str) -> None:
    """Call method ``fn`` on every sub-module that has it (depth first)."""
    for _, m in module.named_children():
        if hasattr(m, fn):
            getattr(m, fn)()
        apply_fn(m, fn)


def save_img(img: torch.Tensor, path: str, nrow: int = 10, padding: int = 2, pad_value: float = 0, range=None) -> None:
    from torchvision.utils import save_image
    makedir_exist_ok(os.path.dirname(path))
    kw = {} if range is None else {"normalize": True, "value_range": range}
    save_image(img, path, nrow=nrow, padding=padding, pad_value=pad_value, **kw)
