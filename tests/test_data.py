import types

import numpy as np
import pytest
import torch

from hfl_b200.config.control import make_cfg
from hfl_b200.data.datasets import batchify, fetch_dataset, process_dataset, synth_lm, synth_vision
from hfl_b200.data.split import split_dataset


def _mini_vision(n=2000, classes=10):
    ds = synth_vision("CIFAR10", "train", n)
    return ds


@pytest.mark.parametrize("mode", ["iid", "non-iid-2"])
def test_split_partition_properties(mode):
    torch.manual_seed(0)
    d = {"train": synth_vision("CIFAR10", "train", 2000), "test": synth_vision("CIFAR10", "test", 400)}
    ds, ls = split_dataset(d, 20, mode, 10)
    for split in ("train", "test"):
        allidx = sum((ds[split][u] for u in range(20)), [])
        assert len(allidx) == len(set(allidx))
        assert set(allidx) <= set(range(len(d[split])))
    if mode == "iid":
        assert isinstance(ls, dict) and all(len(ds["train"][u]) == 100 for u in range(20))
    else:
        assert isinstance(ls, list) and all(1 <= len(l) <= 2 for l in ls)
        for u in range(20):
            assert set(d["train"].target[ds["train"][u]].tolist()) <= set(ls[u])
            assert set(d["test"].target[ds["test"][u]].tolist()) <= set(ls[u])


@pytest.mark.parametrize("mode", ["iid", "non-iid-2"])
def test_split_matches_reference_rng(ref, mode):
    m = ref.load()
    ref.set_cfg("CIFAR10", "resnet18", "1_20_0.1_%s_fix_a1_bn_1_1" % mode, classes_size=10)
    d = {"train": synth_vision("CIFAR10", "train", 2000), "test": synth_vision("CIFAR10", "test", 400)}
    rd = {k: types.SimpleNamespace(target=v.target.tolist(), __len__=None) for k, v in d.items()}

    class _DS:
        def __init__(self, t):
            self.target = t

        def __len__(self):
            return len(self.target)
    rd = {k: _DS(v.target.tolist()) for k, v in d.items()}
    torch.manual_seed(3)
    rds, rls = m["data"].split_dataset(rd, 20, mode)
    torch.manual_seed(3)
    ds, ls = split_dataset(d, 20, mode, 10)
    assert rds == ds
    assert rls == ls


def test_lm_batchify_and_iid():
    cfg = make_cfg("WikiText2", "transformer", "1_10_0.5_iid_fix_a1_none_1_0", synthetic=True,
                   synthetic_size={"train": 10000, "test": 2000}, synthetic_vocab=200)
    d = fetch_dataset(cfg)
    process_dataset(cfg, d)
    assert cfg["num_tokens"] == 200
    assert d["train"].token.shape == (100, 100) and d["test"].token.shape == (10, 200)
    torch.manual_seed(0)
    ds, ls = split_dataset(d, 10, "iid")
    assert sorted(sum(ds["train"].values(), [])) == list(range(100))
    for u in range(10):
        assert ls[u] == torch.unique(d["train"].token[ds["train"][u]]).tolist()


def test_emnist_raw_reader(tmp_path):
    """EMNIST idx files are read transposed, `letters` labels become 0-based (reference src/datasets/mnist.py:86-130)."""
    import gzip
    import struct

    import numpy as np

    from hfl_b200.data.datasets import _load_emnist_raw
    g = tmp_path / "gzip"
    g.mkdir()
    img = np.random.randint(0, 255, (5, 28, 28), dtype=np.uint8)
    lab = np.array([1, 2, 3, 26, 5], dtype=np.uint8)
    with gzip.open(g / "emnist-letters-train-images-idx3-ubyte.gz", "wb") as f:
        f.write(struct.pack(">IIII", 2051, 5, 28, 28) + img.tobytes())
    with gzip.open(g / "emnist-letters-train-labels-idx1-ubyte.gz", "wb") as f:
        f.write(struct.pack(">II", 2049, 5) + lab.tobytes())
    x, y = _load_emnist_raw(str(tmp_path), "train", "letters")
    assert x.shape == (5, 28, 28, 1) and (x[..., 0] == img.transpose(0, 2, 1)).all()
    assert y.tolist() == [0, 1, 2, 25, 4]
    assert _load_emnist_raw(str(tmp_path), "test", "letters") is None
