"""How does train-step throughput scale with the number of concurrently replayed slots (streams)?"""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from hfl_b200.config.control import make_cfg
from hfl_b200.fed.engine import Engine, EngineOptions
from hfl_b200.parallel.dist import DistInfo

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = make_cfg("CIFAR10", "resnet18", "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1", synthetic=True,
               synthetic_size={"train": 50000, "test": 1000})
eng = Engine(cfg, DistInfo(0, 1, 0, dev, None), options=EngineOptions(out_dir="/tmp/p", quiet=True, save=False,
                                                                     max_concurrent=16))
eng.setup()
eng.prewarm(16)
out = {"max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default")}
REP = 100
for lv in (4, 2, 0):
    for n in (1, 2, 4, 8, 12, 16):
        pool = eng._slots[lv][:n]
        key = list(pool[0].graphs.keys())[0]
        for s in pool:
            s.step_ctr.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(REP):
            for s in pool:
                with torch.cuda.stream(s.stream):
                    s.graphs[key].replay()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t1 = time.perf_counter() - t0
        out["level%d_n%d" % (lv, n)] = {"us_per_step": t1 / (REP * n) * 1e6, "issue_us_per_step": t_issue / (REP * n) * 1e6}
    # threaded issue with 8 slots
    pool = eng._slots[lv][:8]
    for s in pool:
        s.step_ctr.zero_()

    def drive(s):
        with torch.cuda.stream(s.stream):
            for _ in range(REP):
                s.graphs[key].replay()
    th = [threading.Thread(target=drive, args=(s,)) for s in pool]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    out["level%d_n8_threads" % lv] = {"us_per_step": (time.perf_counter() - t0) / (REP * 8) * 1e6}
print(json.dumps(out, indent=1))
