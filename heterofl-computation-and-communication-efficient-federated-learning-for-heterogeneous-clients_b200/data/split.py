"""Client data partitioning: ``iid`` and ``non-iid-<N>`` (N class-shards per user).

Semantics and RNG call order follow reference ``src/data.py:48-110`` so that a given seed reproduces the same
partition: ``iid`` draws users one after another from the *remaining* pool with ``randperm``; ``non-iid-N`` cuts
every class into ``N*users/classes`` shards, deals the (shuffled) class list out ``N`` per user -- a user that
draws the same class twice keeps ONE shard of it -- and then hands each user a random remaining shard of each of
its classes.  ``label_split`` is a dict for iid and a list of lists for non-iid (both index like ``ls[user]``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .datasets import VisionData


def _labels_of(ds) -> torch.Tensor:
    return ds.target if isinstance(ds, VisionData) else ds.token


def iid(ds, num_users: int) -> Tuple[Dict[int, List[int]], Dict[int, List[int]]]:
    label = _labels_of(ds)
    n = len(ds)
    per_user = int(n / num_users)
    pool = list(range(n))
    data_split, label_split = {}, {}
    for u in range(num_users):
        take = min(len(pool), per_user)
        chosen = torch.tensor(pool)[torch.randperm(len(pool))[:take]].tolist()
        data_split[u] = chosen
        label_split[u] = torch.unique(label[chosen]).tolist()
        pool = list(set(pool) - set(chosen))
    return data_split, label_split


def non_iid(ds: VisionData, num_users: int, shard_per_user: int, classes_size: int,
            label_split: Optional[List[List[int]]] = None):
    label = np.asarray(ds.target)
    by_class: Dict[int, List[int]] = {}
    for i, c in enumerate(label.tolist()):
        by_class.setdefault(c, []).append(i)
    shard_per_class = int(shard_per_user * num_users / classes_size)
    shards: Dict[int, List] = {}
    for c, idx in by_class.items():
        rem = len(idx) % shard_per_class
        body = np.array(idx[:len(idx) - rem]) if rem > 0 else np.array(idx)
        parts = body.reshape((shard_per_class, -1)).tolist()
        for j, extra in enumerate(idx[len(idx) - rem:] if rem > 0 else []):
            parts[j] = np.concatenate([parts[j], [extra]])
        shards[c] = parts
    if label_split is None:
        deck = list(range(classes_size)) * shard_per_class
        deck = torch.tensor(deck)[torch.randperm(len(deck))].tolist()
        label_split = np.array(deck).reshape((num_users, -1)).tolist()
        label_split = [np.unique(row).tolist() for row in label_split]
    data_split: Dict[int, List[int]] = {u: [] for u in range(num_users)}
    for u in range(num_users):
        for c in label_split[u]:
            pick = torch.arange(len(shards[c]))[torch.randperm(len(shards[c]))[0]].item()
            data_split[u].extend(int(v) for v in shards[c].pop(pick))
    return data_split, label_split


def split_dataset(dataset: Dict[str, object], num_users: int, data_split_mode: str, classes_size: int = 0):
    """-> (data_split {'train': {user: idx}, 'test': {...}}, label_split)  (reference ``src/data.py:48-58``)."""
    data_split = {}
    if data_split_mode == "iid":
        data_split["train"], label_split = iid(dataset["train"], num_users)
        data_split["test"], _ = iid(dataset["test"], num_users)
    elif "non-iid" in data_split_mode:
        n = int(data_split_mode.split("-")[-1])
        data_split["train"], label_split = non_iid(dataset["train"], num_users, n, classes_size)
        data_split["test"], _ = non_iid(dataset["test"], num_users, n, classes_size, label_split)
    else:
        raise ValueError("Not valid data split mode")
    return data_split, label_split
