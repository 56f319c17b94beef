import os
import subprocess
import sys

import pytest

from hfl_b200.utils.make import ablation_grid, combinations, control_grid, interpolations, make_script

REF = "/root/reference/src"


def test_grid_sizes():
    assert len(combinations(list("abcde"))) == 26 and len(interpolations(list("abcde"))) == 90
    assert len(control_grid(1, "iid")) == 5 + 26 + 90 and len(control_grid(0, "iid")) == 5


@pytest.mark.parametrize("script,args", [
    ("make.py", ["--run", "train", "--model", "resnet18", "--fed", "1", "--num_gpu", "4", "--round", "8",
                 "--data_split_mode", "non-iid-2"]),
    ("make.py", ["--run", "test", "--model", "conv", "--fed", "0", "--num_gpu", "2", "--world_size", "2",
                 "--num_experiments", "2"]),
    ("make_ablation.py", ["--run", "train", "--model", "transformer", "--data_split_mode", "iid", "--num_gpu", "3"]),
    ("make_ablation.py", ["--run", "train", "--model", "resnet18", "--data_split_mode", "non-iid-2"]),
])
def test_script_identical_to_reference(tmp_path, script, args):
    if not os.path.isdir(REF):
        pytest.skip("reference not present")
    d1, d2 = tmp_path / "ref", tmp_path / "ours"
    d1.mkdir(), d2.mkdir()
    subprocess.run([sys.executable, os.path.join(REF, script)] + args, cwd=d1, check=True, capture_output=True)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, script)] + args, cwd=d2, check=True, capture_output=True)
    f1 = sorted(os.listdir(d1))
    assert f1 == sorted(os.listdir(d2)) and len(f1) == 1
    assert open(d1 / f1[0]).read() == open(d2 / f1[0]).read()


def test_torchrun_launcher():
    _, s = make_script("train", "resnet18", 1, 8, 8, 0, 1, 1, 1, 0, "iid", launcher="torchrun")
    assert "--nproc-per-node 8" in s and "train_classifier_fed.py" in s
