import os

import numpy as np
import torch

from hfl_b200.data.extra import (BoundingBoxCrop, ClassTree, Compose, Stats, calculate_md5, extract_file,
                                 load_image_folder, load_omniglot, make_classes_counts)


def test_stats_streaming_equals_batch():
    x = torch.randn(500, 3, 8, 8) * torch.tensor([1.0, 2.0, 0.5]).view(1, 3, 1, 1) + torch.tensor([0.1, -1.0, 3.0]).view(1, 3, 1, 1)
    st = Stats(dim=1)
    for s in range(0, 500, 64):
        st.update(x[s:s + 64])
    flat = x.transpose(1, -1).reshape(-1, 3).double()
    assert torch.allclose(st.mean, flat.mean(0), atol=1e-9)
    assert torch.allclose(st.std, flat.std(0), atol=2e-3)
    assert st.n_samples == flat.shape[0]


def test_folder_and_omniglot_readers(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(0)
    for split, n in (("train", 3), ("test", 2)):
        for c in ("cat", "dog"):
            d = tmp_path / "Pets" / "raw" / split / c
            d.mkdir(parents=True)
            for i in range(n):
                Image.fromarray(rng.randint(0, 255, (12, 10, 3), dtype=np.uint8)).save(d / ("%d.png" % i))
    tr = load_image_folder(str(tmp_path / "Pets"), "train", size=8)
    te = load_image_folder(str(tmp_path / "Pets"), "test", size=8)
    assert tr.img.shape == (6, 8, 8, 3) and te.img.shape == (4, 8, 8, 3) and tr.classes_size == 2
    assert tr.target.tolist() == [0, 0, 0, 1, 1, 1]
    for a in ("Greek", "Latin"):
        for ch in ("character01", "character02"):
            d = tmp_path / "Omniglot" / "raw" / "images_background" / a / ch
            d.mkdir(parents=True)
            Image.fromarray(rng.randint(0, 255, (20, 20), dtype=np.uint8)).save(d / "0.png")
    om = load_omniglot(str(tmp_path / "Omniglot"), "train", size=14)
    assert om.img.shape == (4, 14, 14, 1) and om.classes_size == 4


def test_tree_transforms_and_archive_helpers(tmp_path):
    t = ClassTree()
    for path in (["fish", "ray"], ["fish", "shark"], ["trees", "oak_tree"]):
        t.add(path)
    assert t.make_flat_index() == 3 and t.find("shark").flat_index == 1 and t.find("shark").index == [0, 1]
    assert make_classes_counts([0, 1, 1, 2])[1] == 2
    from PIL import Image
    img = Image.fromarray(np.zeros((10, 10, 3), dtype=np.uint8))
    out = Compose([BoundingBoxCrop(), lambda im: im.size])({"img": img, "bbox": torch.tensor([1, 2, 4, 5])})
    assert out["img"] == (4, 5)
    import zipfile
    z = tmp_path / "a.zip"
    with zipfile.ZipFile(z, "w") as f:
        f.writestr("x.txt", "hello")
    extract_file(str(z))
    assert open(tmp_path / "x.txt").read() == "hello" and len(calculate_md5(str(z))) == 32
