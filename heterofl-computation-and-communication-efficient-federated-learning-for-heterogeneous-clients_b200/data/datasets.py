"""Datasets as plain tensors.

The reference wraps every dataset in a ``torch.utils.data.Dataset`` yielding one PIL image at a time and augments
on the CPU (``src/data.py:10-34``, ``src/datasets/*.py``).  On a B200 the whole of CIFAR-10 is 150 MB of uint8 -- it
lives in HBM (or pinned host memory) as ONE tensor ``img[N, H, W, C]`` and batches are gathered + augmented by a
kernel.  Sources, tried in order by :func:`fetch_dataset`:

1. a processed cache written by the reference (``./data/<NAME>/processed/{train,test}.pt`` -- same tuple format,
   ``src/datasets/cifar.py:50-56``), so an existing data directory keeps working;
2. raw files (MNIST idx, CIFAR python pickles, WikiText/PTB token files) under ``./data/<NAME>/raw``;
3. synthetic data of the named shape (``cfg['synthetic']`` -- there is no network on the build box): class-structured
   images / Zipfian token streams with the real datasets' sizes.
"""
from __future__ import annotations

import codecs
import gzip
import os
import pickle
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

DATA_STATS = {  # mean / std used by the reference transforms (src/data.py:15-30)
    "MNIST": ((0.1307,), (0.3081,)), "FashionMNIST": ((0.1307,), (0.3081,)), "EMNIST": ((0.1307,), (0.3081,)),
    "CIFAR10": ((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)),
    "CIFAR100": ((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)),
}
DATA_SIZES = {"MNIST": (60000, 10000, 10, (28, 28, 1)), "FashionMNIST": (60000, 10000, 10, (28, 28, 1)),
              "EMNIST": (60000, 10000, 10, (28, 28, 1)),
              "CIFAR10": (50000, 10000, 10, (32, 32, 3)), "CIFAR100": (50000, 10000, 100, (32, 32, 3))}
LM_SIZES = {"WikiText2": (2088628, 245569, 33278), "PennTreebank": (929589, 82430, 10000),
            "WikiText103": (103227021, 245569, 267735)}


class Vocab:
    """Symbol table with ``<ukn>``=0 and ``<eos>``=1 (reference ``src/datasets/lm.py:9-51``)."""

    def __init__(self):
        self.symbol_to_index = {u"<ukn>": 0, u"<eos>": 1}
        self.index_to_symbol = [u"<ukn>", u"<eos>"]

    def add(self, symbol):
        if symbol not in self.symbol_to_index:
            self.symbol_to_index[symbol] = len(self.index_to_symbol)
            self.index_to_symbol.append(symbol)

    def delete(self, symbol):
        if symbol in self.symbol_to_index:
            self.index_to_symbol.remove(symbol)
            self.symbol_to_index.pop(symbol, None)

    def __len__(self):
        return len(self.index_to_symbol)

    def __getitem__(self, key):
        if isinstance(key, int):
            return self.index_to_symbol[key] if 0 <= key < len(self.index_to_symbol) else u"<ukn>"
        if isinstance(key, str):
            return self.symbol_to_index.get(key, self.symbol_to_index[u"<ukn>"])
        raise ValueError("Not valid data type")

    def __contains__(self, key):
        if isinstance(key, int):
            return 0 <= key < len(self.index_to_symbol)
        if isinstance(key, str):
            return key in self.symbol_to_index
        raise ValueError("Not valid data type")


class SizedVocab(Vocab):
    """Synthetic vocabulary of a given size without materialising 33k python strings up front."""

    def __init__(self, n: int):
        super().__init__()
        for i in range(2, n):
            self.add("tok%d" % i)


@dataclass
class VisionData:
    data_name: str
    split: str
    img: torch.Tensor          # uint8 [N, H, W, C]
    target: torch.Tensor       # int64 [N]
    classes_size: int
    mean: tuple
    std: tuple
    synthetic: bool = False

    def __len__(self):
        return self.img.shape[0]


@dataclass
class LMData:
    data_name: str
    split: str
    token: torch.Tensor        # int64 [T] (raw) or [rows, T/rows] after batchify
    vocab: Vocab
    synthetic: bool = False

    def __len__(self):
        return self.token.shape[0]


def batchify(ds: LMData, batch_size: int) -> LMData:
    """Reshape the token stream to ``batch_size`` rows (reference ``src/utils.py:353-357``)."""
    n = ds.token.numel() // batch_size
    ds.token = ds.token.reshape(-1)[:n * batch_size].reshape(batch_size, -1)
    return ds


# -- raw readers ------------------------------------------------------------------------------------------
def _open(path):
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def _find(folder, name):
    for cand in (name, name + ".gz"):
        p = os.path.join(folder, cand)
        if os.path.exists(p):
            return p
    return None


def read_idx_images(path) -> np.ndarray:
    with _open(path) as f:
        data = f.read()
    assert int(codecs.encode(data[:4], "hex"), 16) == 2051, "bad idx3 magic"
    n, r, c = (int(codecs.encode(data[o:o + 4], "hex"), 16) for o in (4, 8, 12))
    return np.frombuffer(data, dtype=np.uint8, offset=16).reshape(n, r, c)


def read_idx_labels(path) -> np.ndarray:
    with _open(path) as f:
        data = f.read()
    assert int(codecs.encode(data[:4], "hex"), 16) == 2049, "bad idx1 magic"
    n = int(codecs.encode(data[4:8], "hex"), 16)
    return np.frombuffer(data, dtype=np.uint8, offset=8).reshape(n).astype(np.int64)


def _read_cifar_batches(folder, names):
    imgs, labels = [], []
    for nm in names:
        with open(os.path.join(folder, nm), "rb") as f:
            e = pickle.load(f, encoding="latin1")
        imgs.append(e["data"])
        labels.extend(e["labels"] if "labels" in e else e["fine_labels"])
    img = np.vstack(imgs).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)
    return np.ascontiguousarray(img), np.asarray(labels, dtype=np.int64)


# Where the reference downloads each corpus from (src/datasets/{mnist,cifar,lm}.py `file` lists): (url, md5).
DOWNLOADS = {
    "MNIST": [("http://yann.lecun.com/exdb/mnist/train-images-idx3-ubyte.gz", "f68b3c2dcbeaaa9fbdd348bbdeb94873"),
              ("http://yann.lecun.com/exdb/mnist/train-labels-idx1-ubyte.gz", "d53e105ee54ea40749a09fcbcd1e9432"),
              ("http://yann.lecun.com/exdb/mnist/t10k-images-idx3-ubyte.gz", "9fb629c4189551a2d022fa330f9573f3"),
              ("http://yann.lecun.com/exdb/mnist/t10k-labels-idx1-ubyte.gz", "ec29112dd5afa0611ce80d1b7f02629c")],
    "FashionMNIST": [
        ("http://fashion-mnist.s3-website.eu-central-1.amazonaws.com/train-images-idx3-ubyte.gz",
         "8d4fb7e6c68d591d4c3dfef9ec88bf0d"),
        ("http://fashion-mnist.s3-website.eu-central-1.amazonaws.com/train-labels-idx1-ubyte.gz",
         "25c81989df183df01b3e8a0aad5dffbe"),
        ("http://fashion-mnist.s3-website.eu-central-1.amazonaws.com/t10k-images-idx3-ubyte.gz",
         "bef4ecab320f06d8554ea6380940ec79"),
        ("http://fashion-mnist.s3-website.eu-central-1.amazonaws.com/t10k-labels-idx1-ubyte.gz",
         "bb300cfdad3c16e7a12a480ee83cd310")],
    "EMNIST": [("http://www.itl.nist.gov/iaui/vip/cs_links/EMNIST/gzip.zip", "58c8d27c78d21e728a6bc7b3cc06412e")],
    "CIFAR10": [("https://www.cs.toronto.edu/~kriz/cifar-10-python.tar.gz", "c58f30108f718f92721af3b95e74349a")],
    "CIFAR100": [("https://www.cs.toronto.edu/~kriz/cifar-100-python.tar.gz", "eb9058c3a382ffc7106e4002c42a8d85")],
    "PennTreebank": [("https://raw.githubusercontent.com/wojzaremba/lstm/master/data/ptb.train.txt", None),
                     ("https://raw.githubusercontent.com/wojzaremba/lstm/master/data/ptb.valid.txt", None),
                     ("https://raw.githubusercontent.com/wojzaremba/lstm/master/data/ptb.test.txt", None)],
    "WikiText2": [("https://s3.amazonaws.com/research.metamind.io/wikitext/wikitext-2-v1.zip", None)],
    "WikiText103": [("https://s3.amazonaws.com/research.metamind.io/wikitext/wikitext-103-v1.zip", None)],
}
# EMNIST subsets (reference src/datasets/mnist.py:86-130); `subset` selects which idx files are read
EMNIST_SUBSETS = ("byclass", "bymerge", "balanced", "letters", "digits", "mnist")


def download_dataset(data_name: str, root: str) -> bool:
    """Fetch and unpack the raw files of ``data_name`` into ``root/raw`` (the reference's ``download()``:
    ``src/datasets/mnist.py:59-66``, ``cifar.py``, ``lm.py``).  Returns False when the network is unreachable."""
    import socket
    from .extra import download_url, extract_file
    raw = os.path.join(root, "raw")
    old_to = socket.getdefaulttimeout()
    socket.setdefaulttimeout(float(os.environ.get("HFL_DOWNLOAD_TIMEOUT", "15")))
    try:
        for url, md5 in DOWNLOADS.get(data_name, []):
            fn = os.path.basename(url)
            download_url(url, raw, fn, md5)
            if fn.endswith((".gz", ".zip", ".tgz", ".tar")):
                extract_file(os.path.join(raw, fn))
        return True
    except Exception as e:          # no egress (this sandbox), DNS failure, moved URL ...
        import warnings
        warnings.warn("could not download %s: %s" % (data_name, e))
        return False
    finally:
        socket.setdefaulttimeout(old_to)


def _load_emnist_raw(raw, split, subset):
    """EMNIST idx files (``raw/gzip/emnist-<subset>-{train,test}-{images-idx3,labels-idx1}-ubyte[.gz]``); images are
    stored transposed (reference ``src/datasets/mnist.py:113-118``)."""
    sub = subset if subset in EMNIST_SUBSETS else "balanced"
    folder = os.path.join(raw, "gzip")
    pi = _find(folder, "emnist-%s-%s-images-idx3-ubyte" % (sub, split))
    pl = _find(folder, "emnist-%s-%s-labels-idx1-ubyte" % (sub, split))
    if not (pi and pl):
        return None
    img = np.ascontiguousarray(np.transpose(read_idx_images(pi), (0, 2, 1)))
    lab = read_idx_labels(pl)
    if sub == "letters":          # labels 1..26 in the files; class indices are 0-based
        lab = lab - lab.min()
    return img[..., None], lab


def _load_vision_raw(data_name, root, split):
    raw = os.path.join(root, "raw")
    if data_name == "EMNIST":
        return None               # needs the subset name: handled by fetch_dataset through _load_emnist_raw
    if data_name in ("MNIST", "FashionMNIST"):
        pre = "train" if split == "train" else "t10k"
        pi, pl = _find(raw, pre + "-images-idx3-ubyte"), _find(raw, pre + "-labels-idx1-ubyte")
        if pi and pl:
            return read_idx_images(pi)[..., None], read_idx_labels(pl)
    if data_name == "CIFAR10":
        d = os.path.join(raw, "cifar-10-batches-py")
        if os.path.isdir(d):
            names = ["data_batch_%d" % i for i in range(1, 6)] if split == "train" else ["test_batch"]
            return _read_cifar_batches(d, names)
    if data_name == "CIFAR100":
        d = os.path.join(raw, "cifar-100-python")
        if os.path.isdir(d):
            return _read_cifar_batches(d, [split])
    return None


def _load_vision_processed(root, split, subset):
    p = os.path.join(root, "processed", "%s.pt" % split)
    if not os.path.exists(p):
        return None
    try:
        img, target = torch.load(p, map_location="cpu", weights_only=False)
        img = np.asarray(img)
        if img.ndim == 3:
            img = img[..., None]
        return img, np.asarray(target[subset], dtype=np.int64)
    except Exception:
        return None


# -- synthetic generators ---------------------------------------------------------------------------------
def synth_vision(data_name: str, split: str, n: Optional[int] = None, seed: int = 0) -> VisionData:
    """Class-structured random images: each class has a smooth prototype; samples = prototype + noise.  Learnable
    (convergence smoke tests), with exactly balanced labels like CIFAR/MNIST shards expect."""
    ntr, nte, C, (H, W, Ch) = DATA_SIZES[data_name]
    N = n if n is not None else (ntr if split == "train" else nte)
    g = torch.Generator().manual_seed(1234 + seed)            # prototypes shared by train and test
    proto = torch.rand(C, 8, 8, Ch, generator=g)
    proto = torch.nn.functional.interpolate(proto.permute(0, 3, 1, 2), size=(H, W), mode="bilinear",
                                            align_corners=False).permute(0, 2, 3, 1)
    g2 = torch.Generator().manual_seed(99 + seed + (0 if split == "train" else 1))
    target = torch.arange(N) % C
    target = target[torch.randperm(N, generator=g2)]
    img = torch.empty(N, H, W, Ch, dtype=torch.uint8)
    chunk = 10000
    for s in range(0, N, chunk):
        t = target[s:s + chunk]
        x = proto[t] * 0.6 + 0.4 * torch.rand(t.numel(), H, W, Ch, generator=g2)
        img[s:s + chunk] = (x * 255).clamp_(0, 255).to(torch.uint8)
    mean, std = DATA_STATS[data_name]
    return VisionData(data_name, split, img, target.long(), C, mean, std, synthetic=True)


def synth_lm(data_name: str, split: str, n: Optional[int] = None, vocab: Optional[int] = None,
             seed: int = 0) -> LMData:
    """Zipf-distributed token stream with first-order structure (next token depends on the previous one)."""
    ntr, nte, V = LM_SIZES[data_name]
    V = vocab if vocab is not None else V
    N = n if n is not None else (ntr if split == "train" else nte)
    g = torch.Generator().manual_seed(4321 + seed + (0 if split == "train" else 1))
    ranks = torch.arange(1, V + 1, dtype=torch.float64)
    p = (1.0 / ranks) / (1.0 / ranks).sum()
    base = torch.multinomial(p.float(), N, replacement=True, generator=g)
    # bigram-ish structure: with prob 1/2 a token is a deterministic function of its predecessor
    follow = (base.roll(1) * 7 + 3) % V
    coin = torch.rand(N, generator=g) < 0.5
    tok = torch.where(coin, follow, base)
    tok[::23] = 1  # <eos>
    return LMData(data_name, split, tok.long(), SizedVocab(V), synthetic=True)


# -- public API -------------------------------------------------------------------------------------------
def fetch_dataset(cfg: Dict, root: str = "./data") -> Dict[str, object]:
    """``{'train': ..., 'test': ...}`` for ``cfg['data_name']`` (reference ``src/data.py:10-34``)."""
    name, subset = cfg["data_name"], cfg.get("subset", "label")
    droot = os.path.join(root, name)
    out = {}
    synth = cfg.get("synthetic", "auto")
    sizes = cfg.get("synthetic_size") or {}
    if not sizes and os.environ.get("HFL_SYNTH_SIZES"):          # "train,test" override for smoke tests
        a, b = os.environ["HFL_SYNTH_SIZES"].split(",")
        sizes = {"train": int(a), "test": int(b)}
    origin = {}
    for split in ("train", "test"):
        if name in DATA_SIZES:
            got = None
            if synth is not True:
                def _raw():
                    if name == "EMNIST":
                        return _load_emnist_raw(os.path.join(droot, "raw"), split, subset)
                    return _load_vision_raw(name, droot, split)
                got = _load_vision_processed(droot, split, subset) or _raw()
                if got is None and cfg.get("download", True) and download_dataset(name, droot):
                    got = _raw()
            if got is not None:
                img, tgt = got
                mean, std = DATA_STATS[name]
                out[split] = VisionData(name, split, torch.from_numpy(np.ascontiguousarray(img)),
                                        torch.from_numpy(tgt).long(), int(tgt.max()) + 1, mean, std)
            elif synth in (True, "auto"):
                out[split] = synth_vision(name, split, sizes.get(split), cfg.get("init_seed", 0))
                origin[split] = "synthetic"
            else:
                raise FileNotFoundError("no data for %s under %s and synthetic data disabled" % (name, droot))
        elif name in LM_SIZES:
            tokp = os.path.join(droot, "processed", "%s.pt" % split)
            if synth is not True and os.path.exists(tokp):
                from ..ckpt.compat import load as cload
                out[split] = LMData(name, split, cload(tokp).long(), cload(os.path.join(droot, "processed", "meta.pt")))
            elif synth is not True and (_lm_raw(name, droot, split) is not None or (
                    cfg.get("download", True) and download_dataset(name, droot) and _lm_raw(name, droot, split) is not None)):
                out[split] = _lm_raw(name, droot, split)
            elif synth in (True, "auto"):
                out[split] = synth_lm(name, split, sizes.get(split), cfg.get("synthetic_vocab"),
                                      cfg.get("init_seed", 0))
                origin[split] = "synthetic"
            else:
                raise FileNotFoundError("no data for %s under %s and synthetic data disabled" % (name, droot))
        else:
            raise ValueError("Not valid dataset name")
        origin.setdefault(split, "real")
    kinds = set(origin.values())
    if len(kinds) > 1:
        raise FileNotFoundError("refusing to mix real and synthetic splits of %s (%s); pass --synthetic 1 or provide "
                                "both splits under %s" % (name, origin, droot))
    cfg["synthetic_data"] = kinds == {"synthetic"}
    if cfg["synthetic_data"] and synth == "auto":
        import warnings
        warnings.warn("%s not found under %s: training on SYNTHETIC data of the same shape (results are not %s "
                      "results; recorded as cfg['synthetic_data']=True in checkpoints)" % (name, droot, name),
                      stacklevel=2)
    return out


_LM_FILES = {"WikiText2": ("wikitext-2", "wiki.%s.tokens"), "WikiText103": ("wikitext-103", "wiki.%s.tokens"),
             "PennTreebank": ("", "ptb.%s.txt")}
_LM_VOCAB_CACHE: Dict[str, Vocab] = {}


def _lm_raw(name, droot, split):
    sub, pat = _LM_FILES[name]
    folder = os.path.join(droot, "raw", sub)
    tr = os.path.join(folder, pat % "train")
    if not os.path.exists(tr):
        return None
    if name not in _LM_VOCAB_CACHE:
        v = Vocab()
        files = [tr] + ([os.path.join(folder, pat % "valid")] if name == "PennTreebank" else [])
        for fp in files:
            with open(fp, "r", encoding="utf-8") as f:
                for line in f:
                    for sym in line.split() + [u"<eos>"]:
                        v.add(sym)
        _LM_VOCAB_CACHE[name] = v
    v = _LM_VOCAB_CACHE[name]
    toks: List[int] = []
    with open(os.path.join(folder, pat % split), "r", encoding="utf-8") as f:
        for line in f:
            toks.extend(v[sym] for sym in line.split() + [u"<eos>"])
    return LMData(name, split, torch.tensor(toks, dtype=torch.long), v)


def process_dataset(cfg: Dict, dataset: Dict[str, object]) -> None:
    """Set ``classes_size`` (vision) or ``vocab``/``num_tokens`` (+ batchify) -- reference ``src/utils.py:100-110``."""
    tr = dataset["train"]
    if isinstance(tr, VisionData):
        cfg["classes_size"] = tr.classes_size
    else:
        cfg["vocab"] = tr.vocab
        cfg["num_tokens"] = len(tr.vocab)
        for split in dataset:
            dataset[split] = batchify(dataset[split], cfg["batch_size"][split])
