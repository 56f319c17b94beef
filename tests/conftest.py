import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_SRC = "/root/reference/src"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class _RefOracle:
    """Imports the unmodified reference modules (read-only) to act as an oracle.  Skips when absent."""

    def __init__(self):
        self.mods = None

    def load(self):
        if self.mods is not None:
            return self.mods
        if not os.path.isdir(REFERENCE_SRC):
            pytest.skip("reference sources not present")
        if "anytree" not in sys.modules:
            try:
                import anytree  # noqa: F401
            except Exception:
                sys.modules["anytree"] = types.ModuleType("anytree")
        cwd = os.getcwd()
        os.chdir(REFERENCE_SRC)
        sys.path.insert(0, REFERENCE_SRC)
        try:
            import config as rconfig
            import fed as rfed
            import models as rmodels
            import utils as rutils
            import data as rdata
        finally:
            os.chdir(cwd)
            sys.path.remove(REFERENCE_SRC)
        self.mods = dict(config=rconfig, fed=rfed, models=rmodels, utils=rutils, data=rdata)
        return self.mods

    def reset_cfg(self):
        """Restore the reference's global cfg to the pristine config.yml contents (its entry-point modules build an
        argparse flag per cfg key at import time and choke on keys derived by earlier process_control() calls)."""
        import yaml
        m = self.load()
        with open(os.path.join(REFERENCE_SRC, "config.yml"), "r") as f:
            fresh = yaml.load(f, Loader=yaml.FullLoader)
        cfg = m["config"].cfg
        cfg.clear()
        cfg.update(fresh)
        return cfg

    def set_cfg(self, data_name, model_name, control_name, **extra):
        m = self.load()
        cfg = m["config"].cfg
        keys = list(cfg["control"].keys()) if cfg.get("control") else \
            ["fed", "num_users", "frac", "data_split_mode", "model_split_mode", "model_mode", "norm", "scale", "mask"]
        cfg["data_name"], cfg["model_name"] = data_name, model_name
        cfg["control"] = {k: v for k, v in zip(keys, control_name.split("_"))}
        cfg["control_name"] = control_name
        cfg["device"] = "cpu"
        m["utils"].process_control()
        for k, v in extra.items():
            cfg[k] = v
        return cfg


@pytest.fixture(scope="session")
def ref():
    return _RefOracle()
