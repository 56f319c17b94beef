"""Server-side evaluation: static-BatchNorm statistics pass + Local/Global test.

Reference ``stats()`` (``src/train_classifier_fed.py:127-138``) pushes the whole train set through a
``track_running_stats=True`` copy of the global model in train-mode batches of 10 (5 000 forwards per round on
CIFAR-10) and ``test()`` (``:141-169``) runs 100 per-user loaders plus one global loader.  Here

* the sBN pass evaluates G ghost batches per launch (:class:`NormCtx` ``stats`` mode), is sharded over ranks and
  merged with one tiny all-reduce of ``[sum mean, sum var]`` vectors;
* the test pass computes the logits of every test image ONCE: *Global* metrics use them as is, *Local* metrics apply
  each sample's owner's label mask (zero-fill) to the same logits -- identical numbers, half the forwards.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from ..data.datasets import VisionData
from ..data.pipeline import VisionAugment, lm_windows
from ..models import Model
from ..models.fused_resnet import FusedResNet
from ..ops import cuda as cu
from ..ops.nn import NormCtx, masked_cross_entropy
from ..parallel.dist import DistInfo, all_reduce_sum


class Evaluator:
    def __init__(self, model: Model, cfg: Dict, info: DistInfo, compute_dtype=torch.float32,
                 stats_chunk_groups: int = 1000, test_chunk: int = 2500):
        self.model, self.cfg, self.info = model, cfg, info
        self.device, self.dtype = info.device, compute_dtype
        self.stats_chunk_groups, self.test_chunk = stats_chunk_groups, test_chunk
        self._aug = {}
        self.use_graph = bool(cfg.get("eval_graph", True))
        self._sbn_g = None
        self._test_g = None
        self.fused = None
        if self.device.type == "cuda" and compute_dtype == torch.bfloat16 and cu.available() \
                and cfg.get("fused", True) and FusedResNet.eligible(model.spec, train_only=False):
            self.fused = FusedResNet(model.spec, model.scale_rate)
        # masked-LM evaluation: the encoder forward on the fused LM kernels (models/fused_transformer.py)
        self.fused_lm = None
        if self.device.type == "cuda" and compute_dtype == torch.bfloat16 and cu.available() and cfg.get("fused", True) \
                and model.is_lm and os.environ.get("HFL_LM_EVAL_FUSED", "1") != "0":
            from ..models.fused_transformer import FusedTransformer
            if FusedTransformer.eligible(model.spec, int(model.layout.entries[
                    "transformer_embedding.embedding.weight"].shape[1])):
                self.fused_lm = FusedTransformer(model.spec, model.scale_rate, mask_rate=float(cfg.get("mask_rate", 0.15)))
                self._lm_seed = torch.tensor([int(cfg.get("seed", 0)) * 7919 + 17], dtype=torch.int64, device=self.device)
                self._lm_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        # resident-kernel programs (models/mega_resnet.py): the whole sBN pass / test pass of this rank is ONE launch
        self.mega = None
        self._mega_sbn = None
        self._mega_test = None
        if self.fused is not None and os.environ.get("HFL_MEGA_EVAL", os.environ.get("HFL_MEGA", "0")) == "1":
            from .mega_trainer import MegaSlot
            if MegaSlot.eligible(model, cfg):
                from ..models.mega_resnet import MegaResNet
                self.mega_grid = int(os.environ.get("HFL_MEGA_EVAL_GRID", "0")) or int(cu.C().mega_max_grid())
                self.mega = MegaResNet(model, self.device, self.mega_grid)
                self.mega_stats_groups = int(os.environ.get("HFL_MEGA_STATS_GROUPS", "200"))
                self.mega_test_chunk = int(os.environ.get("HFL_MEGA_TEST_CHUNK", "2000"))

    def _params(self, global_arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        wc = global_arena if self.dtype == torch.float32 else global_arena.to(self.dtype)
        return self.model.layout.views(wc)

    def _augment(self, ds: VisionData, train: bool) -> VisionAugment:
        key = (ds.split, train)
        if key not in self._aug:
            self._aug[key] = VisionAugment(ds, train, self.device, self.dtype)
        return self._aug[key]

    # -- sBN ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sbn_stats(self, global_arena: torch.Tensor, ds: VisionData, imgs_dev: torch.Tensor,
                  generator: Optional[torch.Generator] = None) -> Dict:
        """Running statistics of every BN layer = average over train-mode batches of ``batch_size['train']``."""
        bn = self.model.bn_layers()
        if not bn:
            return {}
        B = int(self.cfg["batch_size"]["train"])
        N = imgs_dev.shape[0]
        if self.mega is not None and N % B == 0:
            plan = self._chunk_plan(N // B, self.mega_stats_groups)
            if plan is not None:
                return self._sbn_mega(global_arena, ds, imgs_dev, generator, B, plan)
        if self.fused is not None and self.use_graph and N % B == 0:
            plan = self._chunk_plan(N // B, self.stats_chunk_groups)
            if plan is not None:
                return self._sbn_graphed(global_arena, ds, imgs_dev, generator, B, plan)
        if not imgs_dev.is_cuda and self.device.type == "cuda":
            imgs_dev = imgs_dev.to(self.device, non_blocking=True)
        P = self._params(global_arena)
        perm = torch.randperm(N, generator=generator)
        n_full = N // B
        aug = self._augment(ds, True)
        params = aug.draw(N, generator)
        perm_d = perm.to(self.device)
        ctx = NormCtx("stats", group=B)
        G = self.stats_chunk_groups
        chunks = [(g0, min(g0 + G, n_full)) for g0 in range(0, n_full, G)]
        for ci, (g0, g1) in enumerate(chunks):
            if ci % self.info.world != self.info.rank:
                continue
            idx = perm_d[g0 * B:g1 * B]
            x = aug(imgs_dev, idx, params[g0 * B:g1 * B].contiguous())
            if self.fused is not None:
                self.fused.forward_stats(P, x, ctx)
            else:
                self.model.forward(P, {"img": x}, ctx, None, want_score="raw")
            ctx.n_groups += g1 - g0
        if N % B and self.info.rank == 0:
            idx = perm_d[n_full * B:]
            x = aug(imgs_dev, idx, params[n_full * B:].contiguous())
            ctx.group = int(idx.numel())
            self.model.forward(P, {"img": x}, ctx, None, want_score="raw")   # odd-sized tail: generic path
            ctx.n_groups += 1
        # merge across ranks: one flat vector [sum_mean..., sum_var..., n_groups]
        widths = [w for _, w in bn]
        flat = torch.zeros(2 * sum(widths) + 1, dtype=torch.float32, device=self.device)
        off = 0
        for name, w in bn:
            if name in ctx.acc:
                flat[off:off + w] = ctx.acc[name][0]
                flat[off + w:off + 2 * w] = ctx.acc[name][1]
            off += 2 * w
        flat[-1] = ctx.n_groups
        all_reduce_sum(self.info, flat)
        n = flat[-1].clamp(min=1.0)
        running, off = {}, 0
        for name, w in bn:
            running[name] = (flat[off:off + w] / n, flat[off + w:off + 2 * w] / n)
            off += 2 * w
        return running

    # -- graph-captured evaluation (fused path) ----------------------------------------------------------
    def _chunk_plan(self, n_units: int, max_per_chunk: int):
        """(units per chunk, chunks per rank) with every rank getting the same number of equal chunks, or None."""
        w = self.info.world
        for per_rank in range(1, 4096):
            nch = per_rank * w
            if n_units % nch == 0 and n_units // nch <= max_per_chunk:
                return n_units // nch, per_rank
            if nch > n_units:
                break
        return None

    def _merge_stats(self, bn, acc, n_groups: int) -> Dict:
        widths = [w for _, w in bn]
        flat = torch.zeros(2 * sum(widths) + 1, dtype=torch.float32, device=self.device)
        off = 0
        for name, w in bn:
            if name in acc:
                flat[off:off + w] = acc[name][0]
                flat[off + w:off + 2 * w] = acc[name][1]
            off += 2 * w
        flat[-1] = n_groups
        all_reduce_sum(self.info, flat)
        n = flat[-1].clamp(min=1.0)
        running, off = {}, 0
        for name, w in bn:
            running[name] = (flat[off:off + w] / n, flat[off + w:off + 2 * w] / n)
            off += 2 * w
        return running

    def _capture(self, fn):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            fn()                                   # warm-up (allocations, cuDNN plans)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        k0 = cu.launch_count()
        with torch.cuda.graph(g, stream=s):
            fn()
        return g, cu.launch_count() - k0

    def _sbn_graphed(self, global_arena, ds, imgs_src, generator, B, plan) -> Dict:
        """sBN pass as ``chunks_per_rank`` replays of ONE captured chunk graph (no per-kernel host work, so the pass can
        overlap the next round's training when launched on a side stream)."""
        Gc, per_rank = plan
        bn = self.model.bn_layers()
        N = imgs_src.shape[0]
        st = self._sbn_g
        if st is None or st["key"] != (N, B, Gc, per_rank):
            wc = torch.empty(self.model.layout.size, dtype=self.dtype, device=self.device)
            st = {"key": (N, B, Gc, per_rank), "wc": wc, "P": self.model.layout.views(wc),
                  "imgs": torch.empty(imgs_src.shape, dtype=torch.uint8, device=self.device),
                  "perm": torch.zeros(per_rank, Gc * B, dtype=torch.int64, device=self.device),
                  "par": torch.zeros(per_rank, Gc * B, 3, dtype=torch.int32, device=self.device),
                  "row": torch.zeros(1, dtype=torch.int64, device=self.device),
                  "ctx": NormCtx("stats", group=B), "graph": None}
            for name, w in bn:
                st["ctx"].acc[name] = (torch.zeros(w, dtype=torch.float32, device=self.device),
                                       torch.zeros(w, dtype=torch.float32, device=self.device))
            self._sbn_g = st
        aug = self._augment(ds, True)
        perm = torch.randperm(N, generator=generator)
        params = aug.draw(N, generator)
        w_, r_ = self.info.world, self.info.rank
        rows = [ci for ci in range(per_rank * w_) if ci % w_ == r_]
        pv = perm.view(per_rank * w_, Gc * B)[rows].contiguous()
        st["perm"].copy_(pv.to(self.device, non_blocking=True))
        st["par"].copy_(params.view(per_rank * w_, Gc * B, 3)[torch.as_tensor(rows, device=params.device)])
        st["imgs"].copy_(imgs_src, non_blocking=True)
        st["wc"].copy_(global_arena)
        st["row"].zero_()
        for a, b in st["ctx"].acc.values():
            a.zero_()
            b.zero_()
        if st["graph"] is None:
            def chunk():
                idx = st["perm"].index_select(0, st["row"]).view(-1)
                par = st["par"].index_select(0, st["row"]).view(-1, 3)
                x = aug(st["imgs"], idx, par)
                self.fused.forward_stats(st["P"], x, st["ctx"])
                st["row"].add_(1)
            snap = {k: (a.clone(), b.clone()) for k, (a, b) in st["ctx"].acc.items()}
            st["graph"], st["kernels"] = self._capture(chunk)
            st["row"].zero_()
            for k, (a, b) in snap.items():
                st["ctx"].acc[k][0].copy_(a)
                st["ctx"].acc[k][1].copy_(b)
        for _ in range(per_rank):
            st["graph"].replay()
            cu.launch_count_add(st["kernels"])
        return self._merge_stats(bn, st["ctx"].acc, per_rank * Gc)


    # -- resident-kernel evaluation (one launch per pass) -----------------------------------------------------
    def _mega_state(self, key, imgs_src, steps: int, Nc: int, aug) -> Dict:
        n = self.model.layout.size
        dev = self.device
        return {"key": key, "w": torch.empty(n, dtype=torch.float32, device=dev),
                "wc": torch.empty(n, dtype=torch.bfloat16, device=dev),
                "src": torch.empty(imgs_src.shape, dtype=torch.uint8, device=dev),
                "idx": torch.zeros(steps, Nc, dtype=torch.int64, device=dev),
                "par": torch.zeros(steps, Nc, 3, dtype=torch.int32, device=dev),
                "nvalid": torch.full((steps,), Nc, dtype=torch.int32, device=dev),
                "mean": aug.mean.reshape(-1).float().contiguous(), "istd": aug.inv_std.reshape(-1).float().contiguous(),
                "ctl": torch.zeros(16, dtype=torch.int32, device=dev),
                "jctl": torch.zeros(16, dtype=torch.int32, device=dev)}

    def _sbn_mega(self, global_arena, ds, imgs_src, generator, B, plan) -> Dict:
        """sBN pass: ``per_rank`` chunks of ``Gc`` ghost batches, all inside ONE launch of the phase program (tcgen05
        convolutions with per-ghost-batch statistics in their epilogue; running sums accumulated on the device)."""
        Gc, per_rank = plan
        bn = self.model.bn_layers()
        N = imgs_src.shape[0]
        aug = self._augment(ds, True)
        st = self._mega_sbn
        if st is None or st["key"] != (N, B, Gc, per_rank):
            st = self._mega_state((N, B, Gc, per_rank), imgs_src, per_rank, Gc * B, aug)
            st["prog"] = self.mega.build_forward(st, Gc * B, "stats", group_n=B, augment=aug.augment)
            st["prog_dev"] = st["prog"].finalize(self.device)
            self._mega_sbn = st
        perm = torch.randperm(N, generator=generator)
        params = aug.draw(N, generator)
        w_, r_ = self.info.world, self.info.rank
        rows = [ci for ci in range(per_rank * w_) if ci % w_ == r_]
        st["idx"].copy_(perm.view(per_rank * w_, Gc * B)[rows].contiguous().to(self.device, non_blocking=True))
        st["par"].copy_(params.view(per_rank * w_, Gc * B, 3)[torch.as_tensor(rows, device=params.device)])
        st["src"].copy_(imgs_src, non_blocking=True)
        st["w"].copy_(global_arena)
        st["wc"].copy_(global_arena)
        for a, b in st["acc"].values():
            a.zero_()
            b.zero_()
        from ..models import mega_resnet as mr
        mr.launch([(st["prog_dev"], st["nvalid"], st["jctl"], per_rank)], len(st["prog"]), self.mega_grid, st["ctl"])
        acc = {name: (st["acc"][name][0][:w], st["acc"][name][1][:w]) for name, w in bn}
        return self._merge_stats(bn, acc, per_rank * Gc)

    def _test_mega(self, global_arena, running, ds, imgs_src, labels_src, owner_dev, user_masks_dev, plan):
        """Logits of this rank's share of the test set in ONE launch; Global metrics use them as they are, Local
        metrics apply each sample's owner's label mask to the same logits."""
        Nc, per_rank = plan
        N = imgs_src.shape[0]
        bn = self.model.bn_layers()
        aug = self._augment(ds, False)
        st = self._mega_test
        if st is None or st["key"] != (N, Nc, per_rank):
            st = self._mega_state((N, Nc, per_rank), imgs_src, per_rank, Nc, aug)
            st["prog"] = self.mega.build_forward(st, Nc, "eval", steps=per_rank)
            st["prog_dev"] = st["prog"].finalize(self.device)
            w_, r_ = self.info.world, self.info.rank
            rows = [ci for ci in range(per_rank * w_) if ci % w_ == r_]
            st["idx"].copy_(torch.arange(N).view(per_rank * w_, Nc)[rows].to(self.device))
            st["P"] = self.model.layout.views(st["wc"])
            self._mega_test = st
        st["src"].copy_(imgs_src, non_blocking=True)
        st["w"].copy_(global_arena)
        st["wc"].copy_(global_arena)
        for name, (a, b) in self.fused.eval_affine(st["P"], running).items():
            st["affine"][name][0][:a.numel()].copy_(a)
            st["affine"][name][1][:b.numel()].copy_(b)
        from ..models import mega_resnet as mr
        mr.launch([(st["prog_dev"], None, st["jctl"], per_rank)], len(st["prog"]), self.mega_grid, st["ctl"])
        idx = st["idx"].view(-1)
        z = st["logits"]
        y = labels_src.to(self.device, non_blocking=True).index_select(0, idx)
        acc = torch.zeros(6, dtype=torch.float64, device=self.device)
        acc[3] = F.cross_entropy(z, y, reduction="sum").double()
        acc[4] = (z.argmax(1) == y).sum()
        acc[5] = idx.numel()
        own = owner_dev.index_select(0, idx)
        sel = own >= 0
        use_mask = bool(self.cfg["mask"]) and user_masks_dev is not None
        zl = z.masked_fill(~user_masks_dev[own.clamp(min=0)].bool(), 0.0) if use_mask else z
        ce = F.cross_entropy(zl, y, reduction="none")
        acc[0] = (ce * sel).sum().double()
        acc[1] = ((zl.argmax(1) == y) & sel).sum()
        acc[2] = sel.sum()
        all_reduce_sum(self.info, acc)
        return ("vision", acc, True)

    # -- test --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def test_vision(self, global_arena: torch.Tensor, running: Dict, ds: VisionData, imgs_dev: torch.Tensor,
                    labels_dev: torch.Tensor, owner_dev: Optional[torch.Tensor],
                    user_masks_dev: Optional[torch.Tensor], sync: bool = True):
        """``owner_dev[i]`` = user whose test split holds sample i (-1: none); ``user_masks_dev [U, classes]``."""
        if self.mega is not None and owner_dev is not None:
            plan = self._chunk_plan(imgs_dev.shape[0], self.mega_test_chunk)
            if plan is not None:
                pending = self._test_mega(global_arena, running, ds, imgs_dev, labels_dev, owner_dev, user_masks_dev,
                                          plan)
                return self.finish(pending) if sync else pending
        if self.fused is not None and self.use_graph and owner_dev is not None:
            plan = self._chunk_plan(imgs_dev.shape[0], self.test_chunk)
            if plan is not None:
                pending = self._test_graphed(global_arena, running, ds, imgs_dev, labels_dev, owner_dev,
                                             user_masks_dev, plan)
                return self.finish(pending) if sync else pending
        if not imgs_dev.is_cuda and self.device.type == "cuda":
            imgs_dev, labels_dev = imgs_dev.to(self.device, non_blocking=True), labels_dev.to(self.device, non_blocking=True)
        P = self._params(global_arena)
        ctx = NormCtx("eval", running=running)
        ab = self.fused.eval_affine(P, running) if self.fused is not None else None
        aug = self._augment(ds, False)
        N = imgs_dev.shape[0]
        acc = torch.zeros(6, dtype=torch.float64, device=self.device)
        chunks = [(s, min(s + self.test_chunk, N)) for s in range(0, N, self.test_chunk)]
        for ci, (s, e) in enumerate(chunks):
            if ci % self.info.world != self.info.rank:
                continue
            idx = torch.arange(s, e, device=self.device)
            x = aug(imgs_dev, idx, None)
            y = labels_dev[s:e]
            if ab is not None:
                z = self.fused.forward_eval(P, x, ab)
            else:
                z = self.model.forward(P, {"img": x}, ctx, None, want_score="raw")["logits"]
            acc[3] += F.cross_entropy(z, y, reduction="sum").double()
            acc[4] += (z.argmax(1) == y).sum()
            acc[5] += e - s
            if owner_dev is not None:
                own = owner_dev[s:e]
                sel = own >= 0
                if bool(self.cfg["mask"]) and user_masks_dev is not None:
                    m = user_masks_dev[own.clamp(min=0)].bool()
                    zl = z.masked_fill(~m, 0.0)
                else:
                    zl = z
                ce = F.cross_entropy(zl, y, reduction="none")
                acc[0] += (ce * sel).sum().double()
                acc[1] += ((zl.argmax(1) == y) & sel).sum()
                acc[2] += sel.sum()
        all_reduce_sum(self.info, acc)
        pending = ("vision", acc, owner_dev is not None)
        return self.finish(pending) if sync else pending

    def _test_graphed(self, global_arena, running, ds, imgs_src, labels_src, owner_dev, user_masks_dev, plan):
        Nc, per_rank = plan
        N = imgs_src.shape[0]
        st = self._test_g
        bn = self.model.bn_layers()
        if st is None or st["key"] != (N, Nc, per_rank):
            wc = torch.empty(self.model.layout.size, dtype=self.dtype, device=self.device)
            st = {"key": (N, Nc, per_rank), "wc": wc, "P": self.model.layout.views(wc),
                  "imgs": torch.empty(imgs_src.shape, dtype=torch.uint8, device=self.device),
                  "labels": torch.empty(N, dtype=torch.int64, device=self.device),
                  "idx": torch.zeros(per_rank, Nc, dtype=torch.int64, device=self.device),
                  "row": torch.zeros(1, dtype=torch.int64, device=self.device),
                  "acc": torch.zeros(6, dtype=torch.float64, device=self.device),
                  "owner": torch.empty_like(owner_dev),
                  "masks": torch.empty_like(user_masks_dev) if user_masks_dev is not None else None,
                  "ab": {name: (torch.zeros(w, dtype=torch.float32, device=self.device),
                                torch.zeros(w, dtype=torch.float32, device=self.device)) for name, w in bn},
                  "graph": None}
            w_, r_ = self.info.world, self.info.rank
            rows = [ci for ci in range(per_rank * w_) if ci % w_ == r_]
            st["idx"].copy_(torch.arange(N).view(per_rank * w_, Nc)[rows].to(self.device))
            self._test_g = st
        aug = self._augment(ds, False)
        st["imgs"].copy_(imgs_src, non_blocking=True)
        st["labels"].copy_(labels_src, non_blocking=True)
        st["owner"].copy_(owner_dev, non_blocking=True)      # the graph only ever reads evaluator-owned buffers
        if st["masks"] is not None:
            st["masks"].copy_(user_masks_dev, non_blocking=True)
        st["wc"].copy_(global_arena)
        for name, (a, b) in self.fused.eval_affine(st["P"], running).items():
            st["ab"][name][0].copy_(a)
            st["ab"][name][1].copy_(b)
        st["row"].zero_()
        st["acc"].zero_()
        use_mask = bool(self.cfg["mask"]) and st["masks"] is not None
        if st["graph"] is None:
            def chunk():
                idx = st["idx"].index_select(0, st["row"]).view(-1)
                x = aug(st["imgs"], idx, None)
                y = st["labels"].index_select(0, idx)
                z = self.fused.forward_eval(st["P"], x, st["ab"])
                acc = st["acc"]
                acc[3] += F.cross_entropy(z, y, reduction="sum").double()
                acc[4] += (z.argmax(1) == y).sum()
                acc[5] += Nc
                own = st["owner"].index_select(0, idx)
                sel = own >= 0
                zl = z.masked_fill(~st["masks"][own.clamp(min=0)].bool(), 0.0) if use_mask else z
                ce = F.cross_entropy(zl, y, reduction="none")
                acc[0] += (ce * sel).sum().double()
                acc[1] += ((zl.argmax(1) == y) & sel).sum()
                acc[2] += sel.sum()
                st["row"].add_(1)
            st["graph"], st["kernels"] = self._capture(chunk)
            st["row"].zero_()
            st["acc"].zero_()
        for _ in range(per_rank):
            st["graph"].replay()
            cu.launch_count_add(st["kernels"])
        acc = st["acc"].clone()
        all_reduce_sum(self.info, acc)
        return ("vision", acc, True)

    @staticmethod
    def finish(pending) -> Dict[str, float]:
        """Host read of an evaluation's device accumulator (the only synchronising step)."""
        kind, acc, flag = pending
        a = acc.tolist()
        if kind == "lm":
            return {"Global-Loss": a[0] / max(a[2], 1), "Global-Perplexity": a[1] / max(a[2], 1), "_n_global": a[2]}
        out = {"Global-Loss": a[3] / max(a[5], 1), "Global-Accuracy": 100.0 * a[4] / max(a[5], 1)}
        if flag:
            out["Local-Loss"] = a[0] / max(a[2], 1)
            out["Local-Accuracy"] = 100.0 * a[1] / max(a[2], 1)
        out["_n_local"], out["_n_global"] = a[2], a[5]
        return out

    @torch.no_grad()
    def test_lm(self, global_arena: torch.Tensor, rows_dev: torch.Tensor, sync: bool = True):
        """Global masked-LM loss / perplexity over bptt windows of the batchified test stream; every window weighs
        ``rows`` (reference ``src/train_transformer_fed.py:127-143``)."""
        P = self._params(global_arena)
        ctx = NormCtx("eval")
        r, T = rows_dev.shape
        bptt = int(self.cfg["bptt"])
        acc = torch.zeros(3, dtype=torch.float64, device=self.device)
        wins = lm_windows(T, bptt)
        mine = [w for wi, w in enumerate(wins) if wi % self.info.world == self.info.rank]
        full = [w for w in mine if w[1] == bptt]
        WB = 8                                   # windows per forward: the encoder runs once on [r*WB, bptt] tokens
        for b0 in range(0, len(full), WB):
            grp = full[b0:b0 + WB]
            tok = torch.cat([rows_dev[:, s0:s0 + bptt] for s0, _ in grp], dim=0).contiguous()      # [len*r, bptt]
            # (asynchronous evaluation on the helper thread: functional model, see fed/engine.py:_async_eval_ok;
            # HFL_LM_EVAL_FUSED_ASYNC=1 is the diagnosis switch)
            if self.fused_lm is not None and (sync or os.environ.get("HFL_LM_EVAL_FUSED_ASYNC") == "1"):
                # encoder on the fused kernels (fresh [MASK] draw per forward: device-side counter), then ONE vocabulary GEMM
                self._lm_step.add_(1)
                d = self.fused_lm.forward_hidden_eval(P, tok, self._lm_seed, self._lm_step)
                from ..ops.linear import tc_linear
                z = tc_linear(d, P["decoder.linear2.weight"], P["decoder.linear2.bias"], 1.0)
            else:
                z = self.model.forward(P, {"label": tok}, ctx, None, want_score="raw")["logits"]    # [len*r*bptt, V]
            z = z.view(len(grp), r * bptt, -1)
            tgt = tok.view(len(grp), r * bptt)
            for wi in range(len(grp)):           # per-window mean CE (every window weighs r, reference semantics)
                lf = masked_cross_entropy(z[wi], tgt[wi], None).double()
                acc[0] += lf * r
                acc[1] += torch.exp(lf) * r
                acc[2] += r
        for s0, l in [w for w in mine if w[1] != bptt]:
            out = self.model.forward(P, {"label": rows_dev[:, s0:s0 + l].contiguous()}, ctx, None, want_score=False)
            lf = out["loss"].double()
            acc[0] += lf * r
            acc[1] += torch.exp(lf) * r
            acc[2] += r
        all_reduce_sum(self.info, acc)
        pending = ("lm", acc, False)
        return self.finish(pending) if sync else pending
