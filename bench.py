#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): federated rounds/s and local-training samples/s on
``CIFAR10 resnet18 1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1`` (synthetic CIFAR-10-shaped data, random init).

One *step* = one complete communication round exactly as the reference runs it each global epoch: sample 10 of 100
clients, distribute width-sliced sub-networks, 5 local epochs of batch-10 SGD per client, heterogeneous FedAvg, the
static-BatchNorm statistics pass over the 50 000-image train set, and the Local + Global test over the 10 000-image
test set.  (Per-round checkpoint writing is excluded from both arms.)

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N ...             # unmodified reference from baseline/_ref

For N > 1 launch with torchrun (one rank per GPU).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# more hardware work queues than the default 8: a round keeps up to 16 client slots + eval in flight on their own streams
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONTROL = "1_100_0.1_non-iid-2_dynamic_a1-b1-c1-d1-e1_bn_1_1"
CONFIGS = {
    "cifar_dynamic": ("CIFAR10", "resnet18", CONTROL),
    "cifar_homog": ("CIFAR10", "resnet18", "1_8_1.0_iid_fix_a1_gn_0_0"),
    "cifar_stress": ("CIFAR10", "resnet18", "1_800_0.1_non-iid-2_fix_a2-b8_bn_1_1"),
    "wikitext": ("WikiText2", "transformer", "1_100_0.1_iid_fix_a5-b5_none_1_0"),
    "mnist": ("MNIST", "conv", "1_8_1.0_iid_fix_a1_bn_1_1"),
}


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu, self.rows, self.stop = gpu_index, [], threading.Event()
        self.th = None

    def _loop(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.th is not None:
            self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                          ("sw_power_cap", 7)):
            if any(len(r) >= 8 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_ours(args):
    import torch
    from hfl_b200.config.control import make_cfg
    from hfl_b200.fed.engine import Engine, EngineOptions
    from hfl_b200.ops import cuda as cu
    from hfl_b200.parallel.dist import all_reduce_max, barrier, init_distributed, shutdown

    info = init_distributed()
    assert info.device.type == "cuda", "bench.py needs a GPU"
    data, model, control = CONFIGS[args.config]
    sizes = {"train": args.train_size, "test": args.test_size} if args.train_size else None
    cfg = make_cfg(data, model, control, synthetic=True, **({"synthetic_size": sizes} if sizes else {}))
    if args.local_epochs:
        cfg["num_epochs"]["local"] = args.local_epochs
    if args.no_fused:
        cfg["fused"] = False
    eng = Engine(cfg, info, options=EngineOptions(out_dir="/tmp/hfl_bench", quiet=True, save=False,
                                                  compute_dtype=args.dtype, max_concurrent=args.max_concurrent,
                                                  agg_backend=args.agg_backend,
                                                  async_eval=(True if args.async_eval else (False if args.sync_eval else None)),
                                                  threaded_issue=not args.single_thread_issue))
    eng.setup()
    t0 = time.time()
    eng.prewarm()
    prewarm_s = time.time() - t0

    def timed(n_rounds, first_epoch):
        """K rounds bracketed by barrier + synchronize on both sides, CUDA events, max over ranks."""
        barrier(info)
        torch.cuda.synchronize()
        k0 = cu.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        samples, last = 0.0, None
        for k in range(n_rounds):
            last = eng.run_round(first_epoch + k, evaluate=not args.no_eval)
            samples += last["train_n"]
        ev = eng.collect_eval()            # the last round's evaluation finishes INSIDE the timed region
        if ev is not None:
            last["test"] = ev["test"]
        e1.record()
        torch.cuda.synchronize()
        barrier(info)
        ms = torch.tensor([e0.elapsed_time(e1)], device=info.device, dtype=torch.float64)
        all_reduce_max(info, ms)
        return float(ms.item()), samples, cu.launch_count() - k0, last

    ep = 1
    for _ in range(args.warmup):
        eng.run_round(ep, evaluate=not args.no_eval)
        ep += 1
    eng.collect_eval()
    # the end-to-end arm below re-runs EXACTLY these rounds (same sampled clients, same starting W_g): snapshot the
    # sampling RNG, the global model and the LR schedule position here
    snap = {"rng": torch.get_rng_state(), "arena": eng.global_arena.clone(), "running": dict(eng.running),
            "sched": (eng.sched.optimizer.state_dict(), eng.sched.scheduler.state_dict())}
    ep0 = ep
    with ClockSampler(info.local_rank) as cs:
        ms, samples, launches, last = timed(args.steps, ep)
    ep += args.steps
    clocks = cs.summary()
    # end-to-end arm: dataset in pinned host memory, every round's inputs are copied H2D inside the timed region,
    # the round's metrics are read back D2H (run_round already does that)
    eng.set_data_residency(False)
    eng.run_round(ep, evaluate=not args.no_eval)         # one untimed round on the pinned-host data path
    eng.collect_eval()
    torch.set_rng_state(snap["rng"])
    eng.global_arena.copy_(snap["arena"])
    eng.running = dict(snap["running"])
    eng.h2d_bytes = eng.d2h_bytes = 0
    ms_e, samples_e, _, _ = timed(args.steps, ep0)
    h2d, d2h = eng.h2d_bytes / args.steps, eng.d2h_bytes / args.steps
    if info.is_main:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        comm = {}
        try:      # aggregate GB/s vs the NVLink / HBM roofline for the last timed round (BASELINE.json's second metric)
            from hfl_b200.utils.commstats import aggregate_traffic
            lv = [eng.tables.level_id(float(r)) for r in last["rates"]]
            comm = aggregate_traffic([l.size for l in eng.tables.level_layouts], lv, eng.tables.global_layout.size,
                                     info.world, float(last["timings_ms"].get("aggregate", 0.0)),
                                     float(peaks.get("hbm_gbs") or 0.0))
        except Exception as e:      # never let the side metric break the benchmark line
            comm = {"error": repr(e)}
        line = {
            "metric": "federated_train_samples_per_sec", "value": samples / (ms / 1e3), "unit": "samples/s",
            "rounds_per_sec": args.steps / (ms / 1e3), "n_gpus": info.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {"auto": "bf16"}.get(args.dtype, args.dtype), "data": "synthetic",
            "impl": "hfl_b200", "engine": "resident-kernel (HFL_MEGA=1)" if os.environ.get("HFL_MEGA") == "1"
            else "cuda-graph",
            "config": {"model": "%s %s %s" % (data, model, control), "global_batch": cfg["batch_size"]["train"],
                       "seq_len": None, "parallelism": "client-parallel x%d (P2P aggregate)" % info.world,
                       "round": "distribute+local-train+aggregate" + ("" if args.no_eval else "+sBN-stats+test"),
                       "eval_overlap": "sBN+test of round r overlap local training of round r+1 (side stream); the "
                                       "last round's evaluation completes inside the timed region"
                       if eng._async_eval_ok() else "none",
                       "train_samples_per_round": samples / args.steps,
                       "l2": ("inputs_exceed_L2 (every round streams the clients' arenas, optimiser state and the "
                              "%s; nothing reusable across rounds fits the 126 MB L2)" %
                              ("token stream + 77 MB vocabulary tables" if model == "transformer"
                               else "150 MB train set (sBN pass)")),
                       "e2e_rounds": "same rounds as `value` (sampling RNG + W_g restored)",
                       "local_epochs": cfg["num_epochs"]["local"], "prewarm_s": round(prewarm_s, 1)},
            "clocks": clocks,
            "e2e": {"value": samples_e / (ms_e / 1e3), "unit": "samples/s", "rounds_per_sec": args.steps / (ms_e / 1e3),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "aggregate_comm": comm,
            "last_round": {"train": last["train"], "test": last.get("test"), "timings_ms": last["timings_ms"]},
            "measured_peaks": {k: peaks.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")},
        }
        print(json.dumps(line))
    shutdown()


def run_reference(args):
    ref_src = os.path.join(ROOT, "baseline", "_ref", "src")
    if not os.path.isdir(ref_src):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/src missing (run baseline/install_ref.sh; "
                                                              "the reference has no setup.py/pyproject so pip cannot install it)"}))
        return
    from baseline.ref_harness import run as ref_run
    data, model, control = CONFIGS[args.config]
    ref_run(args, data, model, control, ClockSampler)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "hfl_b200", "reference"])
    ap.add_argument("--config", default="cifar_dynamic", choices=sorted(CONFIGS))
    ap.add_argument("--no-eval", action="store_true", help="skip the sBN + test phases (NOT the headline metric)")
    ap.add_argument("--train-size", type=int, default=0)
    ap.add_argument("--test-size", type=int, default=0)
    ap.add_argument("--local-epochs", type=int, default=0)
    ap.add_argument("--max-concurrent", type=int, default=16)
    ap.add_argument("--agg-backend", default="auto")
    ap.add_argument("--single-thread-issue", action="store_true", help="A/B: issue all graph replays from one host thread")
    ap.add_argument("--async-eval", action="store_true",
                    help="force: overlap sBN+test of round r with the training of round r+1 on a side stream (default: "
                         "automatic, on from 4 ranks up; measured: no gain on one GPU, 231 -> 213 ms/round on 8)")
    ap.add_argument("--sync-eval", action="store_true", help="force the evaluation inside its own round (A/B)")
    ap.add_argument("--no-fused", action="store_true", help="A/B: autograd + library-BN path instead of the fused path")
    ap.add_argument("--dtype", default="auto", choices=["auto", "bf16", "fp32"],
                    help="compute dtype of this framework's arm (fp32 = precision-matched to the reference's fp32/TF32 arm: "
                         "functional model + library kernels, no bf16 anywhere)")
    ap.add_argument("--ref-autocast", action="store_true",
                    help="reference arm only: run the reference under torch.autocast(bf16) (precision-matched the other "
                         "way; NOT the stock code path, reported as dtype bf16-autocast)")
    ap.add_argument("--ref-shard-eval", action="store_true",
                    help="reference arm, N > 1 only: also shard the sBN pass and the test loops over the ranks (harness "
                         "code around the reference's model; the stock reference evaluates everything on every rank)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch ourselves under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400),
                   os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
