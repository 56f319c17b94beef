"""Isolated / concurrent replay probe of the captured train step for any bench config (bench/step_probe.py is CIFAR-only):
GPU microseconds per step with ONE slot replaying alone (the chain-bound regime of an 8-GPU round) and per (slot-step) with
K slots of the same level replaying concurrently from one host thread (the 1-GPU regime)."""
import argparse
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import CONFIGS                      # noqa: E402
from hfl_b200.config.control import make_cfg   # noqa: E402
from hfl_b200.fed.engine import Engine, EngineOptions   # noqa: E402
from hfl_b200.parallel.dist import DistInfo   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cifar_dynamic")
ap.add_argument("--slots", type=int, default=4)
ap.add_argument("--replays", type=int, default=100)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
data, model, control = CONFIGS[args.config]
cfg = make_cfg(data, model, control, synthetic=True)
eng = Engine(cfg, DistInfo(0, 1, 0, dev, None), options=EngineOptions(out_dir="/tmp/p", quiet=True, save=False))
eng.setup()
eng.prewarm(args.slots)
out = {}
for lv, pool in sorted(eng._slots.items()):
    s = pool[0]
    key = list(s.graphs.keys())[0]
    g = s.graphs[key]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s.stream):
        s.step_ctr.zero_()
        e0.record()
        for _ in range(args.replays):
            g.replay()
        e1.record()
    torch.cuda.synchronize()
    rec = {"rate": eng.tables.rates[lv], "my_kernels": s.graph_kernels[key],
           "isolated_us_per_step": e0.elapsed_time(e1) * 1000.0 / args.replays}
    grp = pool[:args.slots]
    if len(grp) > 1:
        for t in grp:
            t.step_ctr.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.replays):
            for t in grp:
                with torch.cuda.stream(t.stream):
                    t.graphs[list(t.graphs.keys())[0]].replay()
        torch.cuda.synchronize()
        rec["concurrent%d_us_per_slot_step" % len(grp)] = (time.perf_counter() - t0) / (args.replays * len(grp)) * 1e6
    out["level%d" % lv] = rec
print(json.dumps(out))
