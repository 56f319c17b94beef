import os

import numpy as np

from hfl_b200.ckpt.compat import save
from hfl_b200.utils.logger import Logger
from hfl_b200.utils.process import extract_processed_result, make_df, make_stats, process_result


def _result(path, tag, acc, hist):
    tr, te = Logger(None, tensorboard=False), Logger(None, tensorboard=False)
    te.append({"Global-Accuracy": acc, "Local-Accuracy": acc + 1}, "test", n=10)
    tr.history["test/Global-Accuracy"] = hist
    tr.history["test/Local-Accuracy"] = [h + 1 for h in hist]
    save({"cfg": {}, "epoch": 3, "logger": {"train": tr, "test": te}}, os.path.join(path, tag + ".pt"))


def test_process_pipeline(tmp_path):
    rp = str(tmp_path)
    for lvl, n in (("a", 100), ("e", 10)):
        save({"num_params": n, "num_flops": n * 7, "space": n / 4}, os.path.join(rp, "CIFAR10_resnet18_%s.pt" % lvl))
    exp = ["0", "1"]
    controls = []
    for mode in ("a1", "e1", "a5-e5"):
        for e in exp:
            c = [e, "CIFAR10", "label", "resnet18", "1", "100", "0.1", "iid", "fix", mode, "bn", "1", "1"]
            controls.append(tuple(c))
            _result(rp, "_".join(c), 80.0 + int(e) + (0 if mode == "a1" else -5), [10.0, 50.0, 80.0 + int(e)])
    assert make_stats("_".join(controls[-1]), rp) == (0.55, 55.0, 385.0, 13.75)
    te, th, missing = process_result(controls, exp, rp)
    assert not missing
    node = te["CIFAR10"]["label"]["resnet18"]["1"]["100"]["0.1"]["iid"]["fix"]["a1"]["bn"]["1"]["1"]
    assert node["Global-Accuracy"]["mean"] == 80.5 and abs(node["Global-Accuracy"]["std"] - 0.5) < 1e-12
    assert node["Params"]["mean"] == 100
    hist = th["CIFAR10"]["label"]["resnet18"]["1"]["100"]["0.1"]["iid"]["fix"]["a1"]["bn"]["1"]["1"]["Global-Accuracy"]
    assert np.allclose(hist["mean"], [10.0, 50.0, 80.5])
    ext = {}
    extract_processed_result(ext, te, [])
    df = make_df(ext, rp)
    assert any(k.endswith("a-e") for k in df)
    assert os.path.exists(os.path.join(rp, "result.xlsx")) or os.path.exists(os.path.join(rp, "result.csv"))
