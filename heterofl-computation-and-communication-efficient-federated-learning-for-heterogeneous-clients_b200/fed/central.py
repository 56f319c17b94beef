"""Centralised (non-federated) baselines: ``fed=0`` / ``data_split_mode='none'``.

Reference: ``src/train_classifier.py`` / ``src/train_transformer.py`` -- one model at the global rate, plain SGD epochs
over the whole train set (batch 100), sBN statistics pass + test every epoch, ``nn.DataParallel`` when
``world_size > 1`` (``src/train_classifier.py:65-66``).  Here the same :class:`TrainSlot` machinery trains the single
model; with several ranks the batch is split over ranks and the FLAT gradient arena is all-reduced once per step (one
collective per step instead of DataParallel's per-tensor scatter/gather).  Works for every model family incl. the
Bottleneck ResNets (50/101/152) that cannot be federated.
"""
from __future__ import annotations

import os
import shutil
import time
from typing import Dict, Optional

import torch
import torch.distributed as dist

from ..ckpt import compat as ckpt
from ..config.control import model_tag
from ..data.datasets import LMData, fetch_dataset, process_dataset
from ..data.pipeline import VisionAugment
from ..models import Model
from ..parallel.dist import DistInfo, barrier
from ..plan.spec import build_spec
from ..utils.logger import Logger
from ..utils.optim import make_scheduler_state
from .engine import EngineOptions
from .evaluator import Evaluator
from .trainer import TrainSlot


class CentralTrainer:
    def __init__(self, cfg: Dict, info: Optional[DistInfo] = None, dataset: Optional[Dict] = None,
                 options: Optional[EngineOptions] = None, seed: Optional[int] = None):
        self.cfg = cfg
        self.info = info if info is not None else DistInfo()
        self.opt = options if options is not None else EngineOptions()
        self.device = self.info.device
        self.is_cuda = self.device.type == "cuda"
        self.seed = int(cfg.get("init_seed", 0) if seed is None else seed)
        cfg["model_tag"] = model_tag(cfg, self.seed)
        torch.manual_seed(self.seed)
        self.dtype = {"auto": torch.bfloat16 if self.is_cuda else torch.float32, "bf16": torch.bfloat16,
                      "fp32": torch.float32}[self.opt.compute_dtype]
        self.dataset = dataset if dataset is not None else fetch_dataset(cfg)
        process_dataset(cfg, self.dataset)
        self.is_lm = isinstance(self.dataset["train"], LMData)
        self.spec = build_spec(cfg)
        cl = self.is_cuda
        self.host_model = Model(self.spec, cfg, cfg["global_model_rate"], False)
        self.model = Model(self.spec, cfg, cfg["global_model_rate"], cl)
        arena = self.host_model.convert_arena(self.host_model.new_arena(), self.model).to(self.device)
        if self.info.world > 1:
            cfg["fused"] = False                       # gradients must sit in the flat arena for the all-reduce
        self.slot = TrainSlot(self.model, cfg, self.device, self.dtype,
                              use_graph=False if self.info.world > 1 else (self.opt.use_graph or False))
        self.slot.load_params(arena)
        self.slot.use_mask = False
        if self.info.world > 1:
            self.slot.grad_sync = self._sync_grads
        self.evaluator = Evaluator(self.model, cfg, self.info, self.dtype, self.opt.stats_chunk_groups,
                                   self.opt.test_chunk)
        self.sched = make_scheduler_state(cfg, len(self.spec.params))
        self.logger = Logger(os.path.join(self.opt.out_dir, "runs", "train_%s" % cfg["model_tag"]),
                             tensorboard=self.opt.tensorboard)
        self.epoch = 1
        self.running: Dict = {}
        dev = self.device
        tr, te = self.dataset["train"], self.dataset["test"]
        if self.is_lm:
            self.train_rows, self.test_rows = tr.token.to(dev), te.token.to(dev)
        else:
            self.train_img, self.train_lab = tr.img.to(dev), tr.target.to(dev)
            self.test_img, self.test_lab = te.img.to(dev), te.target.to(dev)
            self.aug_train = VisionAugment(tr, True, dev, self.dtype)
        mode = cfg.get("resume_mode", 0)
        if mode in (1, 2):
            self._resume(mode)

    @property
    def global_arena(self) -> torch.Tensor:
        return self.slot.w

    def _sync_grads(self) -> None:
        dist.all_reduce(self.slot.g)
        self.slot.g.div_(self.info.world)

    # ------------------------------------------------------------------------------------------------------
    def run_epoch(self, epoch: int) -> Dict:
        cfg, info, s = self.cfg, self.info, self.slot
        lr = float(self.sched.lr)
        s.begin_client(lr, torch.ones(s.label_rows, dtype=torch.uint8))
        gen = torch.Generator().manual_seed(self.seed * 1000003 + epoch * 7919)
        B = int(cfg["batch_size"]["train"])
        if self.is_lm:
            rows = self.train_rows
            if info.world > 1:
                per = rows.shape[0] // info.world
                rows = rows[info.rank * per:(info.rank + 1) * per]
            it = s.train_lm(rows, 1, int(cfg["bptt"]))
        else:
            N = self.train_img.shape[0]
            perm = torch.randperm(N, generator=gen)
            if info.world > 1:                      # each rank takes its slice of every global batch
                nb = N // B
                perm = perm[:nb * B].view(nb, info.world, B // info.world)[:, info.rank].reshape(-1)
                B = B // info.world
            idx = perm.to(self.device)
            it = s.train_vision(self.train_img[idx], self.train_lab[idx], self.aug_train, 1, B, gen)
        for _ in it:
            pass
        m = s.meter.buf.clone()
        if info.world > 1:
            dist.all_reduce(m)
        t = m.tolist()
        n = max(t[2], 1.0)
        train = {"Loss": t[0] / n}
        train["Perplexity" if self.is_lm else "Accuracy"] = (t[3] / n) if self.is_lm else 100.0 * t[1] / n
        shared = torch.Generator().manual_seed(self.seed * 1000003 + epoch * 7919 + 104729)
        if self.is_lm:
            te = self.evaluator.test_lm(self.global_arena, self.test_rows)
            test = {"Loss": te["Global-Loss"], "Perplexity": te["Global-Perplexity"]}
        else:
            self.running = self.evaluator.sbn_stats(self.global_arena, self.dataset["train"], self.train_img, shared)
            te = self.evaluator.test_vision(self.global_arena, self.running, self.dataset["test"], self.test_img,
                                            self.test_lab, None, None)
            test = {"Loss": te["Global-Loss"], "Accuracy": te["Global-Accuracy"]}
        return {"epoch": epoch, "lr": lr, "train": train, "train_n": t[2], "test": test, "test_n": te["_n_global"]}

    def fit(self, num_epochs: Optional[int] = None) -> None:
        cfg = self.cfg
        total = cfg["num_epochs"] if not isinstance(cfg["num_epochs"], dict) else cfg["num_epochs"]["global"]
        last = total if num_epochs is None else min(total, self.epoch + num_epochs - 1)
        names = cfg["metric_name"]
        while self.epoch <= last:
            ep = self.epoch
            self.logger.safe(True)
            t0 = time.time()
            res = self.run_epoch(ep)
            dt = time.time() - t0
            lg = self.logger
            lg.append(res["train"], "train", n=max(res["train_n"], 1.0))
            lg.append({"info": ["Model: {}".format(cfg["model_tag"]), "Train Epoch: {}(100%)".format(ep),
                                "Learning rate: {}".format(res["lr"]), "Epoch time: {:.3f}s".format(dt),
                                "Experiment Finished Time: {:.0f}s".format((total - ep) * dt)]}, "train", mean=False)
            quiet = self.opt.quiet or not self.info.is_main
            lg.write("train", names["train"], quiet=quiet)
            lg.append(res["test"], "test", n=max(res["test_n"], 1.0))
            lg.append({"info": ["Model: {}".format(cfg["model_tag"]), "Test Epoch: {}(100%)".format(ep)]}, "test",
                      mean=False)
            lg.write("test", names["test"], quiet=quiet)
            self.sched.step(metric=lg.mean.get("train/%s" % cfg["pivot_metric"]))
            lg.safe(False)
            if self.opt.save and self.info.is_main:
                self.save_checkpoint(ep + 1)
            lg.reset()
            self.epoch = ep + 1
            barrier(self.info)
        self.logger.safe(False)

    # ------------------------------------------------------------------------------------------------------
    def state_dict(self):
        return self.model.layout.state_dict(self.global_arena.detach().float().cpu())

    def load_state_dict(self, sd, strict: bool = True) -> None:
        host = self.global_arena.detach().float().cpu().clone()
        self.model.layout.load_state_dict(host, sd, strict=strict)
        self.slot.load_params(host.to(self.device))

    def save_checkpoint(self, next_epoch: int) -> str:
        cfg = self.cfg
        path = os.path.join(self.opt.out_dir, "model", "%s_checkpoint.pt" % cfg["model_tag"])
        obj = {"cfg": dict(cfg), "epoch": next_epoch, "model_dict": self.state_dict(),
               "optimizer_dict": self.sched.optimizer.state_dict(),
               "scheduler_dict": self.sched.scheduler.state_dict(), "logger": self.logger}
        ckpt.save(obj, path)
        pm = "test/%s" % cfg["pivot_metric"]
        if pm in self.logger.mean:
            cur = self.logger.mean[pm]
            better = cur > cfg["pivot"] if cfg["pivot_metric"].endswith("Accuracy") else cur < cfg["pivot"]
            if better:
                cfg["pivot"] = cur
                shutil.copy(path, os.path.join(self.opt.out_dir, "model", "%s_best.pt" % cfg["model_tag"]))
        return path

    def _resume(self, mode: int, load_tag: str = "checkpoint") -> None:
        cfg = self.cfg
        path = os.path.join(self.opt.out_dir, "model", "%s_%s.pt" % (cfg["model_tag"], load_tag))
        if not os.path.exists(path):
            if self.info.is_main and not self.opt.quiet:
                print("Not exists model tag: {}, start from scratch".format(cfg["model_tag"]))
            return
        ck = ckpt.load(path)
        self.load_state_dict(ck["model_dict"])
        if mode == 1:
            self.epoch = ck["epoch"]
            self.sched.optimizer.load_state_dict(ck["optimizer_dict"])
            self.sched.scheduler.load_state_dict(ck["scheduler_dict"])
            self.logger = ck["logger"]
            if self.info.is_main and not self.opt.quiet:
                print("Resume from {}".format(self.epoch))
        else:
            self.logger = Logger(os.path.join(self.opt.out_dir, "runs", "%s" % cfg["model_tag"]),
                                 tensorboard=self.opt.tensorboard)
