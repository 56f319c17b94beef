"""Resident training megakernel probes: us/step by level and grid size, per-phase cycle profile (CTA 0), and the
aggregate throughput of 10 concurrently training clients (2 per level) -- CUDA events, one launch = all steps."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hfl_b200.config.control import make_cfg                    # noqa: E402
from hfl_b200.data.datasets import fetch_dataset, process_dataset   # noqa: E402
from hfl_b200.data.pipeline import VisionAugment                # noqa: E402
from hfl_b200.fed.mega_trainer import MegaSlot                  # noqa: E402
from hfl_b200.models import Model                               # noqa: E402
from hfl_b200.ops import cuda as cu                             # noqa: E402
from hfl_b200.plan.spec import build_spec                       # noqa: E402


def run(slot, imgs, labs, aug, epochs, gen):
    for _ in slot.train_vision(imgs, labs, aug, epochs, 10, gen):
        pass


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
    cap = cu.C().mega_max_grid()
    print("max co-resident CTAs:", cap, flush=True)
    grids = {1.0: [32, 64, 148], 0.5: [16, 32, 64, 148], 0.25: [8, 16, 32], 0.125: [4, 8, 16], 0.0625: [2, 4, 8]}
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    gen = torch.Generator().manual_seed(1)
    slots = {}
    for rate in grids:
        t0 = time.time()
        model = Model(spec, cfg, rate, True)
        slot = MegaSlot(model, cfg, dev, 10, grid=grids[rate][0])
        slot.load_params(torch.randn(model.layout.size, device=dev) * 0.05)
        slot.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
        run(slot, imgs, labs, aug, 1, gen)
        torch.cuda.synchronize()
        slot.check()
        slots[rate] = slot
        print("rate %g: slot ready in %.1f s, %d phases, %d barriers" % (rate, time.time() - t0, len(slot.prog),
                                                                      slot.prog.n_barriers()), flush=True)
    if what in ("all", "bar"):
        # pure grid-barrier cost: a program of empty phases
        from hfl_b200.models import mega_resnet as mr
        prog = mr.Program()
        for _ in range(200):
            prog.add(mr.OP_NOP, 0, 1, name="nop")
        pd = prog.finalize(dev)
        ctl = torch.zeros(16, dtype=torch.int32, device=dev)
        jctl = torch.zeros(16, dtype=torch.int32, device=dev)
        for grid in (8, 32, 64, 148):
            mr.launch([(pd, None, jctl, 10)], len(prog), grid, ctl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mr.launch([(pd, None, jctl, 50)], len(prog), grid, ctl)
            e1.record()
            torch.cuda.synchronize()
            print("grid %3d: %.2f us per empty phase + barrier" % (grid, 1e3 * e0.elapsed_time(e1) / (50 * 200)), flush=True)
    if what in ("all", "grid"):
        for rate, slot in slots.items():
            for grid in grids[rate]:
                slot.grid = min(grid, cap)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(slot, imgs, labs, aug, 2, gen)
                e1.record()
                torch.cuda.synchronize()
                slot.check()
                print("rate %-7g grid %3d  %.1f us/step" % (rate, slot.grid, 1000.0 * e0.elapsed_time(e1) / 100), flush=True)
    if what in ("all", "prof"):
        for rate, grid in ((1.0, 64), (0.0625, 4)):
            slot = slots[rate]
            slot.grid = grid
            slot.prof = torch.zeros(2 * len(slot.prog), dtype=torch.int64, device=dev)
            run(slot, imgs, labs, aug, 1, gen)
            torch.cuda.synchronize()
            pr = slot.prof.view(-1, 2).cpu().double() / 50.0 / 1965.0       # us per step at 1965 MHz
            slot.prof = None
            rows = sorted(zip(slot.prog.names, pr[:, 0].tolist(), pr[:, 1].tolist()), key=lambda r: -(r[1] + r[2]))
            print("== per-phase profile rate %g grid %d: total %.1f us/step (work %.1f, barrier %.1f)" % (
                rate, grid, float(pr.sum()), float(pr[:, 0].sum()), float(pr[:, 1].sum())))
            by = {}
            for nm, a, b in zip(slot.prog.names, pr[:, 0].tolist(), pr[:, 1].tolist()):
                k = nm.split(".")[-1] if not nm.startswith("sgd") else nm
                if "dgrad" in nm:
                    k = "dgrad"
                by.setdefault(k, [0, 0.0, 0.0])
                by[k][0] += 1
                by[k][1] += a
                by[k][2] += b
            for k, v in sorted(by.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
                print("   %-12s x%3d  work %7.1f us  barrier %7.1f us" % (k, v[0], v[1], v[2]))
            for nm, a, b in rows[:12]:
                print("   top %-28s work %6.1f barrier %6.1f" % (nm, a, b))
    if what in ("all", "conc"):
        # one wave of 10 clients (2 per level) inside ONE launch: lock-step phases, shared barriers, all SMs
        from hfl_b200.fed.mega_trainer import launch_wave
        twins = []
        for rate in grids:
            model = Model(spec, cfg, rate, True)
            s2 = MegaSlot(model, cfg, dev, 10)
            s2.load_params(torch.randn(model.layout.size, device=dev) * 0.05)
            s2.begin_client(0.05, torch.ones(10, dtype=torch.uint8))
            twins += [slots[rate], s2]
        ctl = torch.zeros(16, dtype=torch.int32, device=dev)
        for grid in (148, 96):
            for rep, epochs in enumerate((1, 5)):
                for s in twins:
                    s.prepare_client(imgs, labs, aug, epochs, 10, gen)
                prof = torch.zeros(2 * len(twins[0].prog), dtype=torch.int64, device=dev) if rep else None
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                launch_wave(twins, min(grid, cap), ctl, prof)
                e1.record()
                torch.cuda.synchronize()
                assert ctl.tolist()[1] == 0, ctl.tolist()
                ms = e0.elapsed_time(e1)
                steps = 50 * epochs
                print("wave of 10 clients (2 per level), grid %d, %d steps each: %.1f ms = %.1f us per lock-step "
                      "(%.1f us per client-step)" % (grid, steps, ms, 1e3 * ms / steps, 1e2 * ms / steps), flush=True)
                if prof is not None:
                    pr = prof.view(-1, 2).cpu().double() / steps / 1965.0
                    by = {}
                    for nm, a, b in zip(twins[0].prog.names, pr[:, 0].tolist(), pr[:, 1].tolist()):
                        k = nm.split(".")[-1] if not nm.startswith("sgd") else nm
                        if "dgrad" in nm:
                            k = "dgrad"
                        by.setdefault(k, [0, 0.0, 0.0])
                        by[k][0] += 1
                        by[k][1] += a
                        by[k][2] += b
                    print("   CTA0 per lock-step: work %.1f us, barrier-wait %.1f us" % (float(pr[:, 0].sum()), float(pr[:, 1].sum())))
                    for k, v in sorted(by.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
                        print("      %-12s x%3d  work %7.1f us  barrier %7.1f us" % (k, v[0], v[1], v[2]))


if __name__ == "__main__":
    main()
