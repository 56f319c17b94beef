"""Micro-probe of the captured train step: graph node count, CPU cost of a replay, GPU time per step in isolation,
and throughput when several slots are driven from one vs several host threads."""
import json
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from hfl_b200.config.control import make_cfg
from hfl_b200.fed.engine import Engine, EngineOptions
from hfl_b200.ops import cuda as cu
from hfl_b200.parallel.dist import DistInfo

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", synthetic=True,
               synthetic_size={"train": 50000, "test": 1000})
eng = Engine(cfg, DistInfo(0, 1, 0, dev, None), options=EngineOptions(out_dir="/tmp/p", quiet=True, save=False))
eng.setup()
eng.prewarm(4)
out = {}
for lv, pool in sorted(eng._slots.items()):
    s = pool[0]
    key = list(s.graphs.keys())[0]
    g = s.graphs[key]
    # isolated: one slot, 100 replays
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s.stream):
        s.step_ctr.zero_()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(100):
            g.replay()
        t_cpu = (time.perf_counter() - t0) / 100
        e1.record()
    torch.cuda.synchronize()
    out["level%d" % lv] = {"rate": eng.tables.rates[lv], "my_kernels": s.graph_kernels[key],
                           "cpu_us_per_replay": t_cpu * 1e6, "gpu_us_per_step_isolated": e0.elapsed_time(e1) * 10}
# concurrency: 4 slots of the narrowest level, single host thread vs 4 host threads
for lv in (max(eng._slots), min(eng._slots)):
    pool = eng._slots[lv][:4]
    key = list(pool[0].graphs.keys())[0]
    for s in pool:
        s.step_ctr.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        for s in pool:
            with torch.cuda.stream(s.stream):
                s.graphs[key].replay()
    torch.cuda.synchronize()
    t1 = time.perf_counter() - t0
    for s in pool:
        s.step_ctr.zero_()

    def drive(s):
        with torch.cuda.stream(s.stream):
            for _ in range(100):
                s.graphs[key].replay()
    th = [threading.Thread(target=drive, args=(s,)) for s in pool]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    t2 = time.perf_counter() - t0
    out["concurrent4_level%d" % lv] = {"one_thread_us_per_step": t1 / 400 * 1e6, "four_threads_us_per_step": t2 / 400 * 1e6}
import os
out["cpu_count"] = os.cpu_count()
print(json.dumps(out, indent=1))
