"""Resident-kernel client training: a client's whole local training is ONE cooperative kernel launch.

Reference ``Local.train`` (``src/train_classifier_fed.py:185-210``): per step ``zero_grad -> forward -> backward ->
clip_grad_norm_(1) -> SGD.step -> 2 x .item()``, ~300 launches.  :class:`MegaSlot` keeps the same arenas as
:class:`~hfl_b200.fed.trainer.TrainSlot` (``w`` fp32 master filled by the distribute kernel and read by the aggregate
kernel, ``wc`` bf16 compute copy, ``mom`` momentum) plus an fp32 gradient arena, and replaces the per-step CUDA graph by
the phase program of ``models/mega_resnet.py``: the host draws the epoch permutations / augmentation parameters
(same RNG calls as ``TrainSlot.train_vision``), uploads them as step tables and launches ``csrc/mega_kernel.cu`` once
with ``nsteps = epochs * steps_per_epoch``.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from ..data.pipeline import VisionAugment
from ..models import Model
from ..models import mega_resnet as mr
from ..models.mega_resnet import CTL_WORDS, PHASE_DT, MegaResNet
from ..ops import cuda as cu
from ..utils.metrics import DeviceMeter

# CTAs per client grid by width (the cooperative grids of concurrently training clients share the 148 SMs x 2 slots)
DEFAULT_GRIDS = {1.0: 64, 0.5: 32, 0.25: 12, 0.125: 6, 0.0625: 4}


def grid_for_rate(rate: float, cap: int) -> int:
    env = os.environ.get("HFL_MEGA_GRID")
    if env:
        return max(1, min(int(env), cap))
    best = min(DEFAULT_GRIDS.keys(), key=lambda r: abs(r - rate))
    return max(1, min(DEFAULT_GRIDS[best], cap))


class MegaSlot:
    is_mega = True

    def __init__(self, model: Model, cfg: Dict, device, batch: int, grid: Optional[int] = None):
        assert MegaSlot.eligible(model, cfg)
        self.model, self.cfg, self.device = model, cfg, torch.device(device)
        self.batch = int(batch)
        n = model.layout.size
        self.w = torch.zeros(n, dtype=torch.float32, device=device)
        self.wc = torch.zeros(n, dtype=torch.bfloat16, device=device)
        self.mom = torch.zeros(n, dtype=torch.float32, device=device)
        self.g = torch.zeros(n, dtype=torch.float32, device=device)
        self.mixed = True
        self.meter = DeviceMeter(device)
        self.cap = int(cu.C().mega_max_grid()) if cu.available() else 296
        self.grid = int(grid) if grid else grid_for_rate(model.rate, self.cap)
        prio = -1 if model.rate >= 0.5 else 0
        self.stream = torch.cuda.Stream(device=device, priority=prio) if self.device.type == "cuda" else None
        self.net = MegaResNet(model, device, self.grid)
        classes = self.net.classes
        self.mask_t = torch.ones(classes, dtype=torch.uint8, device=device)
        self.use_mask = bool(cfg["mask"])
        self.ctl = torch.zeros(CTL_WORDS, dtype=torch.int32, device=device)      # launch-wide block of solo launches
        self.jctl = torch.zeros(CTL_WORDS, dtype=torch.int32, device=device)     # per-job scalars (gradient norm)
        self.pending = None
        self.lr = 0.0
        self._cap_n = 0
        self._cap_steps = 0
        self.st: Dict = {}
        self.prog = None
        self.prog_dev = None
        self.launches = 0

    @staticmethod
    def eligible(model: Model, cfg: Dict) -> bool:
        spec = model.spec
        if not (spec.family == "resnet" and spec.meta.get("norm") == "bn" and spec.meta.get("kind") == "block"):
            return False
        ds = spec.meta["data_shape"]
        H, W = int(ds[1]), int(ds[2])
        # pixel bricks: every feature-map width must divide 128 / 64 (CIFAR 32 -> 16 -> 8 -> 4)
        n_down = sum(1 for b in spec.meta["blocks"] if b["stride"] == 2)
        return H == W and W in (32, 64) and (W >> n_down) >= 2 and int(ds[0]) <= 8

    # ---------------------------------------------------------------------------------------------------------
    def _ensure(self, aug: VisionAugment, max_n: int, max_steps: int, img_shape) -> None:
        if self.prog is not None and self._cap_n >= max_n and self._cap_steps >= max_steps:
            return
        dev = self.device
        self._cap_n, self._cap_steps = max(max_n, self._cap_n), max(max_steps, self._cap_steps)
        st = self.st = {}
        st["w"], st["wc"], st["g"], st["mom"] = self.w, self.wc, self.g, self.mom
        st["src"] = torch.zeros((self._cap_n,) + tuple(img_shape), dtype=torch.uint8, device=dev)
        st["labels"] = torch.zeros(self._cap_n, dtype=torch.int64, device=dev)
        st["idx"] = torch.zeros(self._cap_steps, self.batch, dtype=torch.int64, device=dev)
        st["par"] = torch.zeros(self._cap_steps, self.batch, 3, dtype=torch.int32, device=dev)
        st["nvalid"] = torch.zeros(self._cap_steps, dtype=torch.int32, device=dev)
        st["mean"] = aug.mean.reshape(-1).float().contiguous()
        st["istd"] = aug.inv_std.reshape(-1).float().contiguous()
        st["mask"] = self.mask_t if self.use_mask else None
        st["meter"] = self.meter.buf
        st["lr"], st["momentum"], st["wd"], st["max_norm"] = 0.0, float(self.cfg["momentum"]), \
            float(self.cfg["weight_decay"]), 1.0
        self.prog = self.net.build_train(st, self.batch)
        self._lr_phase = self.prog.names.index("sgd.update")
        self._host = np.stack(self.prog.recs)
        self.prog_dev = self.prog.finalize(dev)
        self._lr_off = self._lr_phase * PHASE_DT.itemsize + PHASE_DT.fields["fv"][1]
        self._lr_dev = -1.0

    def _set_lr(self, lr: float) -> None:
        if lr != self._lr_dev:
            b = torch.tensor([lr], dtype=torch.float32).view(torch.uint8).to(self.device, non_blocking=True)
            self.prog_dev[self._lr_off:self._lr_off + 4].copy_(b, non_blocking=True)
            self._lr_dev = lr

    # -- per-client setup ----------------------------------------------------------------------------------
    def begin_client(self, lr: float, label_mask: torch.Tensor, rng_seed: Optional[int] = None) -> None:
        self.lr = float(lr)
        self.mask_t.copy_(label_mask.to(self.mask_t.device), non_blocking=True)
        self.meter.zero_()

    def load_params(self, arena: torch.Tensor) -> None:
        self.w.copy_(arena)
        self.wc.copy_(self.w)
        self.mom.zero_()
        self.g.zero_()

    def prepare_client(self, imgs_u8: torch.Tensor, labels: torch.Tensor, aug: VisionAugment, epochs: int, batch: int,
                       generator: Optional[torch.Generator] = None, max_n: int = 0, max_epochs: int = 0):
        """Upload the step tables of one client (same RNG draws as ``TrainSlot.train_vision``) and return the job tuple
        ``(program, nvalid table, per-job ctl, nsteps)`` for :func:`mega_resnet.launch`."""
        assert batch == self.batch
        n = imgs_u8.shape[0]
        spe = (n + batch - 1) // batch
        steps = epochs * spe
        max_n = max(n, max_n)
        self._ensure(aug, max_n, max(epochs, max_epochs) * ((max_n + batch - 1) // batch), imgs_u8.shape[1:])
        st = self.st
        perms = torch.stack([torch.randperm(n, generator=generator) for _ in range(epochs)])
        params = aug.draw(epochs * n, generator).view(epochs, n, 3)
        pad = spe * batch - n
        if pad:
            perms = torch.cat([perms, torch.zeros(epochs, pad, dtype=perms.dtype)], dim=1)
            params = torch.cat([params, torch.zeros(epochs, pad, 3, dtype=params.dtype, device=params.device)], dim=1)
        nv = torch.full((epochs, spe), batch, dtype=torch.int32)
        if pad:
            nv[:, -1] = batch - pad
        st["src"][:n].copy_(imgs_u8, non_blocking=True)
        st["labels"][:n].copy_(labels, non_blocking=True)
        st["idx"][:steps].copy_(perms.view(steps, batch), non_blocking=True)
        st["par"][:steps].copy_(params.view(steps, batch, 3), non_blocking=True)
        st["nvalid"][:steps].copy_(nv.view(-1), non_blocking=True)
        self._set_lr(self.lr)
        self.pending = (self.prog_dev, st["nvalid"], self.jctl, steps)
        return self.pending

    def train_vision(self, imgs_u8: torch.Tensor, labels: torch.Tensor, aug: VisionAugment, epochs: int, batch: int,
                     generator: Optional[torch.Generator] = None, max_n: int = 0, max_epochs: int = 0):
        """Generator with a single step: uploads the step tables and launches the resident kernel for this client alone
        (the engine batches the clients of a wave into one launch instead: :func:`launch_wave`)."""
        job = self.prepare_client(imgs_u8, labels, aug, epochs, batch, generator, max_n, max_epochs)
        mr.launch([job], len(self.prog), self.grid, self.ctl, getattr(self, "prof", None))
        self.launches += 1
        yield

    def check(self) -> None:
        """Raise if the last launch aborted on a device-side timeout (host sync)."""
        c = self.ctl.tolist()
        if c[1] != 0:
            nm = self.prog.names[c[8]] if 0 <= c[8] < len(self.prog.names) else "?"
            raise RuntimeError("megakernel aborted (code %d): diag %s phase %s" % (c[1], c[8:14], nm))

    def metrics(self) -> Dict[str, float]:
        return self.meter.result("Local-", lm=False)


def launch_wave(slots, grid: int, ctl: torch.Tensor, prof: Optional[torch.Tensor] = None) -> None:
    """Train the prepared clients of ``slots`` in ONE cooperative launch (lock-step phases, shared grid barriers)."""
    order = sorted(slots, key=lambda s: -s.model.rate)            # heaviest tiles first
    nph = len(order[0].prog)
    assert all(len(s.prog) == nph and s.pending is not None for s in order)
    mr.launch([s.pending for s in order], nph, grid, ctl, prof)
    for s in order:
        s.pending = None
        s.launches += 1
