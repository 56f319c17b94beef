"""Per-step time of the resident training megakernel by level and grid size (CUDA events, 1 launch = `steps` steps)."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hfl_b200.config.control import make_cfg                    # noqa: E402
from hfl_b200.data.datasets import fetch_dataset, process_dataset   # noqa: E402
from hfl_b200.data.pipeline import VisionAugment                # noqa: E402
from hfl_b200.fed.mega_trainer import MegaSlot                  # noqa: E402
from hfl_b200.models import Model                               # noqa: E402
from hfl_b200.ops import cuda as cu                             # noqa: E402
from hfl_b200.plan.spec import build_spec                       # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", synthetic=True,
                   synthetic_size={"train": 2000, "test": 200})
    ds = fetch_dataset(cfg)
    process_dataset(cfg, ds)
    spec = build_spec(cfg)
    aug = VisionAugment(ds["train"], True, dev, torch.bfloat16)
    imgs = ds["train"].img[:500].to(dev)
    labs = ds["train"].target[:500].to(dev)
    print("max co-resident CTAs:", cu.C().mega_max_grid())
    grids = {1.0: [16, 32, 48, 64, 96, 128, 148], 0.5: [8, 16, 32, 64, 128], 0.25: [4, 8, 16, 32], 0.125: [2, 4, 8, 16],
             0.0625: [1, 2, 4, 8]}
    only = [float(a) for a in sys.argv[1:]] or list(grids)
    for rate in only:
        model = Model(spec, cfg, rate, True)
        for grid in grids[rate]:
            slot = MegaSlot(model, cfg, dev, 10, grid=grid)
            slot.load_params(model.new_arena(dev))
            slot.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
            gen = torch.Generator().manual_seed(1)
            for _ in slot.train_vision(imgs, labs, aug, 1, 10, gen):     # warm-up launch (50 steps)
                pass
            torch.cuda.synchronize()
            slot.check()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in slot.train_vision(imgs, labs, aug, 2, 10, gen):     # 100 steps, one launch
                pass
            e1.record()
            torch.cuda.synchronize()
            slot.check()
            ms = e0.elapsed_time(e1)
            print("rate %-7g grid %3d  phases %d  %.1f us/step  loss %.3f" % (
                rate, grid, len(slot.prog), 1000.0 * ms / 100, slot.metrics()["Local-Loss"]), flush=True)
            del slot


if __name__ == "__main__":
    main()
