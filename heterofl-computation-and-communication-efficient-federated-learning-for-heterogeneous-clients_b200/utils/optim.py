"""Optimizer / LR-scheduler factories (reference ``src/utils.py:260-297``).

Client steps use the fused flat-arena SGD kernel (``csrc/optim_kernels.cu``); the *global* optimizer of the reference
only ever supplies the learning rate and its ``state_dict`` to checkpoints.  :class:`SchedulerState` keeps exactly
that: a torch optimizer over N placeholder parameters (N = number of model tensors, so the saved ``optimizer_dict``
loads into the reference's optimizer) plus the configured scheduler.
"""
from __future__ import annotations

import warnings
from typing import Dict, Iterable

import torch
import torch.optim as optim


def make_optimizer(params: Iterable[torch.Tensor], cfg: Dict, lr: float) -> optim.Optimizer:
    name = cfg["optimizer_name"]
    if name == "SGD":
        return optim.SGD(params, lr=lr, momentum=cfg["momentum"], weight_decay=cfg["weight_decay"])
    if name == "RMSprop":
        return optim.RMSprop(params, lr=lr, momentum=cfg["momentum"], weight_decay=cfg["weight_decay"])
    if name == "Adam":
        return optim.Adam(params, lr=lr, weight_decay=cfg["weight_decay"])
    if name == "Adamax":
        return optim.Adamax(params, lr=lr, weight_decay=cfg["weight_decay"])
    raise ValueError("Not valid optimizer name")


def make_scheduler(optimizer: optim.Optimizer, cfg: Dict):
    name = cfg["scheduler_name"]
    sched = optim.lr_scheduler
    if name == "None":
        return sched.MultiStepLR(optimizer, milestones=[65535])
    if name == "StepLR":
        return sched.StepLR(optimizer, step_size=cfg["step_size"], gamma=cfg["factor"])
    if name == "MultiStepLR":
        return sched.MultiStepLR(optimizer, milestones=cfg["milestones"], gamma=cfg["factor"])
    if name == "ExponentialLR":
        return sched.ExponentialLR(optimizer, gamma=0.99)
    if name == "CosineAnnealingLR":
        ne = cfg["num_epochs"]
        return sched.CosineAnnealingLR(optimizer, T_max=ne["global"] if isinstance(ne, dict) else ne,
                                       eta_min=cfg["min_lr"])
    if name == "ReduceLROnPlateau":
        return sched.ReduceLROnPlateau(optimizer, mode="min", factor=cfg["factor"], patience=cfg["patience"],
                                       threshold=cfg["threshold"], threshold_mode="rel", min_lr=cfg["min_lr"])
    if name == "CyclicLR":
        return sched.CyclicLR(optimizer, base_lr=cfg["lr"], max_lr=10 * cfg["lr"])
    raise ValueError("Not valid scheduler name")


class SchedulerState:
    def __init__(self, cfg: Dict, n_params: int):
        self.cfg = cfg
        self.params = [torch.nn.Parameter(torch.zeros(1)) for _ in range(max(n_params, 1))]
        self.optimizer = make_optimizer(self.params, cfg, cfg["lr"])
        self.scheduler = make_scheduler(self.optimizer, cfg)

    @property
    def lr(self) -> float:
        return self.optimizer.param_groups[0]["lr"]

    def step(self, metric=None) -> None:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if self.cfg["scheduler_name"] == "ReduceLROnPlateau":
                self.scheduler.step(metrics=metric)
            else:
                self.scheduler.step()


def make_scheduler_state(cfg: Dict, n_params: int) -> SchedulerState:
    return SchedulerState(cfg, n_params)
