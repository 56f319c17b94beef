"""The long tail of the reference's data layer, tensor-style.

* :class:`Stats`            -- streaming per-channel mean / std with the parallel-variance merge (``src/utils.py:231-257``)
* :class:`ClassTree`        -- label hierarchy without the ``anytree`` dependency (``src/datasets/utils.py:160-190``)
* download / checksum / extract helpers (``src/datasets/utils.py:70-129``) -- usable when a network exists
* folder-backed datasets: :func:`load_image_folder` (``datasets/folder.py``), :func:`load_imagenet`
  (``datasets/imagenet.py``; pre-resized ``raw/<size>/{train,val}`` folders), :func:`load_omniglot`
  (``datasets/omniglot.py``) -- each decodes the images ONCE into a uint8 ``[N, H, W, C]`` tensor (the device-resident
  format of :class:`hfl_b200.data.datasets.VisionData`) instead of re-opening a file per sample per epoch
* dict-aware transforms ``Compose`` / ``CustomTransform`` / ``BoundingBoxCrop`` (``datasets/transforms.py``,
  ``datasets/utils.py:193-211``) for code that still feeds PIL images.
"""
from __future__ import annotations

import gzip
import hashlib
import os
import tarfile
import zipfile
from collections import Counter
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .datasets import VisionData

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif")


# -- statistics ---------------------------------------------------------------------------------------------
class Stats:
    def __init__(self, dim: int):
        self.dim, self.n_samples, self.n_features = dim, 0, None
        self.mean = self.std = None

    def update(self, data: torch.Tensor) -> None:
        data = data.transpose(self.dim, -1).reshape(-1, data.size(self.dim)).double()
        n = data.size(0)
        new_mean = data.mean(dim=0)
        new_std = data.std(dim=0) if n > 1 else torch.zeros_like(new_mean)
        if self.n_samples == 0:
            self.n_samples, self.n_features, self.mean, self.std = n, data.size(1), new_mean, new_std
            return
        m = float(self.n_samples)
        tot = m + n
        mean = m / tot * self.mean + n / tot * new_mean
        self.std = torch.sqrt(m / tot * self.std ** 2 + n / tot * new_std ** 2
                              + m * n / tot ** 2 * (self.mean - new_mean) ** 2)
        self.mean = mean
        self.n_samples += n


def make_stats(ds: VisionData, batch: int = 100, cache: Optional[str] = None) -> Stats:
    """Per-channel statistics of a dataset (reference ``make_stats``; cached under ``./data/stats`` there)."""
    if cache and os.path.exists(cache):
        return torch.load(cache, weights_only=False)
    st = Stats(dim=1)
    for s in range(0, len(ds), batch):
        st.update(ds.img[s:s + batch].permute(0, 3, 1, 2).float() / 255.0)
    if cache:
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        torch.save(st, cache)
    return st


def make_classes_counts(label) -> Counter:
    label = np.asarray(label)
    if label.ndim > 1:
        label = label.sum(axis=tuple(range(1, label.ndim)))
    return Counter(label.tolist())


# -- label tree ---------------------------------------------------------------------------------------------
class ClassTree:
    """Minimal class hierarchy: ``add(['vehicles 1', 'bus'])`` builds paths, leaves get consecutive flat indices."""

    def __init__(self, name: str = "U"):
        self.name, self.children, self.index, self.flat_index = name, [], [], None

    def find(self, name: str) -> Optional["ClassTree"]:
        if self.name == name:
            return self
        for c in self.children:
            r = c.find(name)
            if r is not None:
                return r
        return None

    def add(self, path: Sequence[str]) -> None:
        if not path:
            return
        node = next((c for c in self.children if c.name == path[0]), None)
        if node is None:
            node = ClassTree(path[0])
            node.index = self.index + [len(self.children)]
            self.children.append(node)
        node.add(path[1:])

    def leaves(self) -> List["ClassTree"]:
        if not self.children:
            return [self]
        out: List[ClassTree] = []
        for c in self.children:
            out += c.leaves()
        return out

    def make_flat_index(self, given: Optional[Sequence[str]] = None) -> int:
        size = 0
        for i, leaf in enumerate(self.leaves()):
            leaf.flat_index = given.index(leaf.name) if given else i
            size = max(size, leaf.flat_index + 1)
        return size


# -- download helpers ---------------------------------------------------------------------------------------
def calculate_md5(path: str, chunk_size: int = 1 << 20) -> str:
    md5 = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(chunk_size), b""):
            md5.update(chunk)
    return md5.hexdigest()


def check_integrity(path: str, md5: Optional[str] = None) -> bool:
    return os.path.isfile(path) and (md5 is None or calculate_md5(path) == md5)


def download_url(url: str, root: str, filename: str, md5: Optional[str] = None) -> str:
    from urllib import request
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, filename)
    if check_integrity(path, md5):
        return path
    try:
        request.urlretrieve(url, path)
    except OSError:
        if url.startswith("https"):
            request.urlretrieve(url.replace("https:", "http:"), path)
        else:
            raise
    if not check_integrity(path, md5):
        raise RuntimeError("Not valid downloaded file")
    return path


def extract_file(src: str, dest: Optional[str] = None, delete: bool = False) -> None:
    dest = os.path.dirname(src) if dest is None else dest
    name = os.path.basename(src)
    if name.endswith(".zip"):
        with zipfile.ZipFile(src, "r") as z:
            z.extractall(dest)
    elif name.endswith((".tar", ".tar.gz", ".tgz")):
        with tarfile.open(src) as t:
            t.extractall(dest)
    elif name.endswith(".gz"):
        with gzip.open(src, "rb") as g, open(src[:-3], "wb") as out:
            out.write(g.read())
    if delete:
        os.remove(src)


# -- folder datasets ----------------------------------------------------------------------------------------
def find_classes(folder: str) -> Dict[str, int]:
    classes = sorted(d.name for d in os.scandir(folder) if d.is_dir())
    return {c: i for i, c in enumerate(classes)}


def _scan(folder: str, classes: Dict[str, int]) -> Tuple[List[str], List[int]]:
    paths, labels = [], []
    for c in sorted(classes):
        d = os.path.join(folder, c)
        if not os.path.isdir(d):
            continue
        for r, _, files in sorted(os.walk(d)):
            for fn in sorted(files):
                if fn.lower().endswith(IMG_EXTENSIONS):
                    paths.append(os.path.join(r, fn))
                    labels.append(classes[c])
    return paths, labels


def _decode(paths: Sequence[str], size: Optional[int], mode: str) -> torch.Tensor:
    from PIL import Image
    imgs = []
    for p in paths:
        with Image.open(p) as im:
            im = im.convert(mode)
            if size is not None and im.size != (size, size):
                im = im.resize((size, size), Image.BILINEAR)
            a = np.asarray(im, dtype=np.uint8)
        imgs.append(a[..., None] if a.ndim == 2 else a)
    return torch.from_numpy(np.stack(imgs)) if imgs else torch.zeros(0, size or 1, size or 1, 3, dtype=torch.uint8)


def load_image_folder(root: str, split: str, size: Optional[int] = None, mode: str = "RGB",
                      mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), name: Optional[str] = None) -> VisionData:
    """``root/raw/{train,test}/<class>/*.png`` -> VisionData (classes from the train folder)."""
    raw = os.path.join(root, "raw")
    if not os.path.isdir(raw):
        raise RuntimeError("Dataset not found")
    classes = find_classes(os.path.join(raw, "train"))
    paths, labels = _scan(os.path.join(raw, split), classes)
    C = 1 if mode == "L" else 3
    return VisionData(name or os.path.basename(root.rstrip("/")), split, _decode(paths, size, mode),
                      torch.tensor(labels, dtype=torch.long), len(classes), tuple(mean)[:C], tuple(std)[:C])


def load_imagenet(root: str, split: str, size: int = 32) -> VisionData:
    """Pre-resized ImageNet: ``root/raw/<size>/{train,val}/<wnid>/*.JPEG`` (the reference's ``ImageNet(size=...)``)."""
    base = os.path.join(root, "raw", str(size))
    if not os.path.isdir(base):
        raise RuntimeError("Dataset not found")
    classes = find_classes(os.path.join(base, "train"))
    paths, labels = _scan(os.path.join(base, "train" if split == "train" else "val"), classes)
    return VisionData("ImageNet", split, _decode(paths, size, "RGB"), torch.tensor(labels, dtype=torch.long),
                      len(classes), (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))


def load_omniglot(root: str, split: str, size: int = 28) -> VisionData:
    """``images_background`` (train) / ``images_evaluation`` (test); a class is an (alphabet, character) pair."""
    folder = os.path.join(root, "raw", "images_background" if split == "train" else "images_evaluation")
    if not os.path.isdir(folder):
        raise RuntimeError("Dataset not found")
    paths, labels, cls = [], [], {}
    for alpha in sorted(os.listdir(folder)):
        for ch in sorted(os.listdir(os.path.join(folder, alpha))):
            key = alpha + "/" + ch
            cls.setdefault(key, len(cls))
            d = os.path.join(folder, alpha, ch)
            for fn in sorted(os.listdir(d)):
                if fn.lower().endswith(IMG_EXTENSIONS):
                    paths.append(os.path.join(d, fn))
                    labels.append(cls[key])
    return VisionData("Omniglot", split, _decode(paths, size, "L"), torch.tensor(labels, dtype=torch.long), len(cls),
                      (0.9221,), (0.2681,))


# -- dict-aware transforms ------------------------------------------------------------------------------------
class CustomTransform:
    def __call__(self, inp):
        return inp["img"]

    def __repr__(self):
        return self.__class__.__name__


class BoundingBoxCrop(CustomTransform):
    def __call__(self, inp):
        x, y, w, h = torch.as_tensor(inp["bbox"]).long().tolist()
        return inp["img"].crop((x, y, x + w, y + h))


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, inp):
        for t in self.transforms:
            inp["img"] = t(inp) if isinstance(t, CustomTransform) else t(inp["img"])
        return inp

    def __repr__(self):
        return self.__class__.__name__ + "(" + "".join("\n    {}".format(t) for t in self.transforms) + "\n)"
