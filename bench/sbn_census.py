"""Kernel breakdown of the sBN statistics pass + test pass (torch.profiler, a few chunks)."""
import collections
import json
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from hfl_b200.config.control import make_cfg
from hfl_b200.fed.engine import Engine, EngineOptions
from hfl_b200.parallel.dist import DistInfo

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", synthetic=True,
               synthetic_size={"train": 10240, "test": 5000})
eng = Engine(cfg, DistInfo(0, 1, 0, dev, None), options=EngineOptions(out_dir="/tmp/p", quiet=True, save=False))
eng.setup()
g = torch.Generator().manual_seed(0)
for _ in range(2):
    run = eng.evaluator.sbn_stats(eng.global_arena, eng.dataset["train"], eng.train_img, g)
torch.cuda.synchronize()
out = {}
for name, fn in (("sbn_10240_images", lambda: eng.evaluator.sbn_stats(eng.global_arena, eng.dataset["train"], eng.train_img, g)),
                 ("test_5000_images", lambda: eng.evaluator.test_vision(eng.global_arena, run, eng.dataset["test"], eng.test_img,
                                                                        eng.test_lab, eng.test_owner, eng.user_masks))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    cnt, tim = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            cnt[ev.name[:80]] += 1
            tim[ev.name[:80]] += ev.device_time
    rows = sorted(tim.items(), key=lambda kv: -kv[1])[:14]
    out[name] = {"event_ms": wall, "sum_kernel_ms": sum(tim.values()) / 1e3, "kernels": sum(cnt.values()),
                 "top": [(k, cnt[k], round(v / 1e3, 3)) for k, v in rows]}
print(json.dumps(out, indent=1))
