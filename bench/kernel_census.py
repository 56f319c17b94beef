"""Which kernels make up one captured train step?  (torch.profiler over eager steps of a TrainSlot, per level.)"""
import collections
import json
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from hfl_b200.config.control import make_cfg
from hfl_b200.data.datasets import synth_vision
from hfl_b200.data.pipeline import VisionAugment
from hfl_b200.fed.trainer import TrainSlot
from hfl_b200.models import Model
from hfl_b200.plan.spec import build_spec

dev = torch.device("cuda", 0)
cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", classes_size=10)
spec = build_spec(cfg)
ds = synth_vision("CIFAR10", "train", 100)
img, lab = ds.img.to(dev), ds.target.to(dev)
out = {}
for rate in (1.0, 0.0625):
    aug = VisionAugment(ds, True, dev, torch.bfloat16)
    model = Model(spec, cfg, rate, True)
    slot = TrainSlot(model, cfg, dev, torch.bfloat16, False)
    slot.load_params(model.new_arena(dev))
    slot.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
    g = slot.train_vision(img, lab, aug, 1, 10)
    for _ in range(5):
        next(g)
    print("fused" if slot.fused is not None else "autograd", file=sys.stderr)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        next(g)
        torch.cuda.synchronize()
    cnt = collections.Counter()
    tim = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            cnt[ev.name[:90]] += 1
            tim[ev.name[:90]] += ev.device_time
    rows = sorted(cnt.items(), key=lambda kv: -kv[1])
    out["rate_%g" % rate] = {"total_kernels": sum(cnt.values()), "total_device_us": sum(tim.values()),
                             "by_kernel": [(k, v, round(tim[k], 1)) for k, v in rows]}
print(json.dumps(out, indent=1))
