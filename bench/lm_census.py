"""Kernel census of one masked-LM train step (transformer, WikiText2 shapes: ROWS x 64 tokens, vocab 33 278)."""
import collections
import json
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from hfl_b200.config.control import make_cfg
from hfl_b200.fed.trainer import TrainSlot
from hfl_b200.models import Model
from hfl_b200.plan.spec import build_spec

dev = torch.device("cuda", 0)
V = 33278
cfg = make_cfg("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_0", num_tokens=V, classes_size=V)
spec = build_spec(cfg)
out = {}
ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 1     # federated WikiText2/100 users: ONE row of 64 tokens per step
rows = torch.randint(0, V, (ROWS, 64 * 12), device=dev)
mask = torch.zeros(V + 1, dtype=torch.uint8)      # +1: the [MASK] row, never a target
mask[torch.randperm(V)[:9000]] = 1
for rate in (1.0, 0.5):
    model = Model(spec, cfg, rate, True)
    slot = TrainSlot(model, cfg, dev, torch.bfloat16, False)
    slot.load_params(model.new_arena(dev))
    slot.begin_client(0.05, mask)
    g = slot.train_lm(rows, 4, 64)
    print("fused" if slot.fused is not None else "autograd", file=sys.stderr)
    for _ in range(5):
        next(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6):
        next(g)
    e1.record()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        next(g)
        torch.cuda.synchronize()
    cnt, tim = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            cnt[ev.name[:100]] += 1
            tim[ev.name[:100]] += ev.device_time
    seq = [(ev.name[:60], round(ev.device_time, 1)) for ev in prof.events()
           if ev.device_type == torch.autograd.DeviceType.CUDA]
    rows_ = sorted(cnt.items(), key=lambda kv: -tim[kv[0]])
    out["rate_%g" % rate] = {"eager_ms_per_step": e0.elapsed_time(e1) / 6, "total_kernels": sum(cnt.values()),
                             "total_device_us": sum(tim.values()),
                             "by_kernel": [(k, v, round(tim[k], 1)) for k, v in rows_[:40]], "sequence": seq}
print(json.dumps(out, indent=1))
