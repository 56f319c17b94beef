"""Client-side local training.

Reference ``Local.train`` (``src/train_classifier_fed.py:185-210``, ``src/train_transformer_fed.py:152-183``) builds a
fresh ``nn.Module`` + fresh SGD per client and loops ``zero_grad -> forward -> backward -> clip_grad_norm_(1) ->
step -> .item() metrics`` with ~300 kernel launches and 2 host syncs per step.

Here a :class:`TrainSlot` is a *resident* execution context for one complexity level:

* parameters in a flat fp32 master arena ``w`` (filled directly by the distribute kernel, read directly by the
  aggregate kernel -- possibly by a peer GPU), a bf16 compute copy ``wc`` and a bf16 gradient arena ``g`` whose
  per-tensor views are the autograd leaves, fp32 momentum ``mom``;
* the client's data shard resident on the device; batches are gathered/augmented by one kernel from a device-side
  step counter, so a whole train step (batch assembly, forward, loss+metrics, backward, clip+SGD) is ONE captured
  CUDA graph that is replayed ``epochs * steps`` times without touching the host;
* metrics accumulate on the device and are read back once per client.

Several slots run concurrently on their own streams, which is how a 148-SM B200 is kept busy by clients whose layers
are as narrow as 4 channels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..data.pipeline import VisionAugment, lm_windows
from ..models import Model
from ..models.fused_conv import FusedConvNet
from ..models.fused_resnet import FusedResNet
from ..models.fused_transformer import EMB, FusedTransformer
from ..ops import cuda as cu
from ..ops.nn import NormCtx
from ..utils.metrics import DeviceMeter


def sgd_clip_step_torch(w, wc, g, mom, lr: float, momentum: float, wd: float, max_norm: float) -> None:
    """Flat-arena reference implementation of clip_grad_norm_ + SGD (used on CPU and as the kernel oracle)."""
    gf = g.float()
    total = gf.norm()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    gg = gf * coef + wd * w
    mom.mul_(momentum).add_(gg)
    w.add_(mom, alpha=-lr)
    if wc is not None and wc.data_ptr() != w.data_ptr():
        wc.copy_(w)
    g.zero_()


class TrainSlot:
    def __init__(self, model: Model, cfg: Dict, device, compute_dtype=torch.float32, use_graph: bool = False,
                 w: Optional[torch.Tensor] = None):
        self.model, self.cfg, self.device = model, cfg, torch.device(device)
        self.dtype = compute_dtype
        n = model.layout.size
        self.w = w if w is not None else torch.zeros(n, dtype=torch.float32, device=device)
        self.mixed = compute_dtype != torch.float32
        self.wc = torch.zeros(n, dtype=compute_dtype, device=device) if self.mixed else self.w
        self.g = torch.zeros(n, dtype=compute_dtype, device=device)
        self.mom = torch.zeros(n, dtype=torch.float32, device=device)
        self.P = {}
        gviews = model.layout.views(self.g)
        for k, v in model.layout.views(self.wc).items():
            leaf = v.detach().requires_grad_(True)
            leaf.grad = gviews[k]
            self.P[k] = leaf
        self.meter = DeviceMeter(device)
        self.is_cuda = self.device.type == "cuda"
        self.use_kernels = self.is_cuda and cu.available()
        self.use_graph = bool(use_graph and self.is_cuda)
        # wider sub-networks sit on the round's critical path (their steps are the longest): give their streams
        # scheduling priority so the narrow clients fill the gaps instead of delaying them
        prio = -1 if model.rate >= 0.5 else 0
        self.stream = torch.cuda.Stream(device=device, priority=prio) if self.is_cuda else None
        self.lr_t = torch.zeros(1, dtype=torch.float32, device=device)
        self.scratch = torch.zeros(4, dtype=torch.float32, device=device)
        self.momentum, self.wd = float(cfg["momentum"]), float(cfg["weight_decay"])
        self.max_norm = 1.0
        self.label_rows = max([model.layout.entries[p.name].shape[0] for p in model.spec.params if p.label_rows] + [1])
        self.mask_t = torch.ones(self.label_rows, dtype=torch.uint8, device=device)
        self.use_mask = bool(cfg["mask"])
        # manual-backward fused path (bf16 + kernels): no autograd graph, no gradient arena traffic
        self.fused = None
        self.rng_seed = torch.zeros(1, dtype=torch.int64, device=device)     # dropout / [MASK] draws of the fused LM
        self._clients = 0
        if self.use_kernels and self.mixed and compute_dtype == torch.bfloat16 and cfg.get("fused", True):
            if FusedResNet.eligible(model.spec):
                self.fused = FusedResNet(model.spec, model.scale_rate)
            elif FusedConvNet.eligible(model.spec):
                self.fused = FusedConvNet(model.spec, model.scale_rate)
            elif model.is_lm and FusedTransformer.eligible(model.spec, model.layout.entries[EMB].shape[1]):
                self.fused = FusedTransformer(model.spec, model.scale_rate, cfg.get("mask_rate", 0.15),
                                              cfg["transformer"]["dropout"])
        if self.fused is not None:
            self.Pv = model.layout.views(self.wc)
            self._names = list(model.layout.entries.keys())
            chunk = cu.C().ms_chunk()
            blk = []
            for ti, name in enumerate(self._names):
                blk += [(ti, s0) for s0 in range(0, model.layout.entries[name].numel, chunk)]
            self.blk_tab = torch.tensor(blk, dtype=torch.int32, device=device).reshape(-1, 2)
            self.seg_tab = torch.zeros(len(self._names), 4, dtype=torch.int64, device=device)
            self._grads = None
            self._capturing = False
            self._captured_segs = None
        self.graphs: Dict = {}
        self.graph_kernels: Dict = {}
        self.step_ctr = torch.zeros(1, dtype=torch.int64, device=device)
        self._static: Dict = {}
        self.lr = 0.0

    # -- per-client setup ----------------------------------------------------------------------------------
    def begin_client(self, lr: float, label_mask: torch.Tensor, rng_seed: Optional[int] = None) -> None:
        """Parameters are expected in ``w`` (+``wc``) already (distribute kernel); reset optimiser state + meter."""
        self._clients += 1
        self.rng_seed.fill_(int(rng_seed) if rng_seed is not None else self._clients * 104729 + 12345)
        self.lr = float(lr)
        self.lr_t.fill_(float(lr))
        self.mask_t.copy_(label_mask.to(self.mask_t.device), non_blocking=True)
        self.meter.zero_()
        self.step_ctr.zero_()

    def load_params(self, arena: torch.Tensor) -> None:
        """Host/fallback path: copy a distributed arena in and reset momentum."""
        self.w.copy_(arena)
        if self.mixed:
            self.wc.copy_(self.w)
        self.mom.zero_()
        self.g.zero_()

    # -- one optimisation step -----------------------------------------------------------------------------
    def _seg_rows(self):
        rows = []
        lay = self.model.layout
        for name in self._names:
            g = self._grads[name]
            e = lay.entries[name]
            if g.dim() == 4:
                dense = g.permute(0, 2, 3, 1).is_contiguous() if lay.channels_last else g.is_contiguous()
            else:
                dense = g.is_contiguous()
            if not dense:   # re-lay into the arena's physical order (not expected on the cuDNN NHWC path)
                g = g.contiguous(memory_format=torch.channels_last if (g.dim() == 4 and lay.channels_last)
                                 else torch.contiguous_format)
                self._grads[name] = g
            assert g.numel() == e.numel, name
            ptr = 0 if (name == EMB and self._emb_sparse()) else g.data_ptr()     # 0: applied row-sparsely afterwards
            rows.append((ptr, e.offset, e.numel, 1 if g.dtype == torch.bfloat16 else 0))
        return rows

    def _emb_sparse(self) -> bool:
        f = self.fused
        return bool(getattr(f, "emb_sparse", False) and getattr(f, "emb_src", None) is not None and self.grad_sync is None)

    grad_sync = None   # optional callable run between backward and the optimiser (data-parallel all-reduce)

    def _opt_step(self) -> None:
        if self.grad_sync is not None:
            self.grad_sync()
        if self.fused is not None:
            rows = self._seg_rows()
            if self._capturing:
                self._captured_segs = rows          # pointers are fixed inside the graph pool: table filled after capture
            else:
                self.seg_tab.copy_(torch.tensor(rows, dtype=torch.int64), non_blocking=False)
                self._seg_active = None
            sparse = self._emb_sparse()
            if sparse:      # row-sparse token-table gradient: its norm first, its update after the dense pass
                ge = self._grads[EMB]
                cu.C().emb_sparse_norm(self.fused.emb_src, ge, self.scratch)
            cu.C().multi_sgd_clip_step(self.seg_tab, self.blk_tab, self.w, self.wc, self.mom, self.lr_t,
                                       self.momentum, self.wd, self.max_norm, self.scratch)
            if sparse:
                cu.C().emb_sparse_apply(self.fused.emb_src, ge, int(self.model.layout.entries[EMB].offset), self.w, self.wc,
                                        self.mom, self.lr_t, self.max_norm, self.scratch)
            if not self._capturing:
                self._grads = None
            return
        if self.use_kernels:
            cu.sgd_clip_step(self.w, self.wc if self.mixed else None, self.g, self.mom, self.lr_t, self.momentum,
                             self.wd, self.max_norm, self.scratch)
        else:
            sgd_clip_step_torch(self.w, self.wc, self.g, self.mom, self.lr, self.momentum, self.wd, self.max_norm)

    def _fwd_bwd(self, inp: Dict, n_weight: float, nvalid: Optional[torch.Tensor] = None) -> None:
        mask = self.mask_t if self.use_mask else None
        if self.fused is not None and self.model.is_lm:
            if self.grad_sync is not None:
                self.fused.emb_sparse = False     # dense all-reduce of every gradient: the token table stays dense
            _, self._grads = self.fused.train_step(self.Pv, inp["label"], mask, (self.meter.buf, n_weight),
                                                   self.rng_seed, self.step_ctr)
            return
        if self.fused is not None:
            _, self._grads = self.fused.train_step(self.Pv, inp["img"], inp["label"], mask, (self.meter.buf, n_weight),
                                                   nvalid)
            return
        if self.use_kernels:
            inp = dict(inp)
            inp["_meter"] = (self.meter.buf, n_weight)
        out = self.model.forward(self.P, inp, NormCtx("train"), mask, want_score=not self.use_kernels
                                 and not self.model.is_lm)
        out["loss"].backward()
        if not self.use_kernels:
            self.meter.add(out["loss"], n_weight, out.get("score"), inp["label"] if "score" in out else None)

    # -- vision --------------------------------------------------------------------------------------------
    def train_vision(self, imgs_u8: torch.Tensor, labels: torch.Tensor, aug: VisionAugment, epochs: int, batch: int,
                     generator: Optional[torch.Generator] = None, max_n: int = 0, max_epochs: int = 0):
        """Generator yielding after every optimisation step (lets the engine interleave slots across streams)."""
        n = imgs_u8.shape[0]
        full, rem = n // batch, n % batch
        steps_per_epoch = full + (1 if rem else 0)
        perms = torch.stack([torch.randperm(n, generator=generator) for _ in range(epochs)])
        params = aug.draw(epochs * n, generator).view(epochs, n, 3)
        perms_d = perms.to(self.device, non_blocking=True)
        # the fused ResNet step takes the number of real samples from the device (``nvalid``), so the partial last
        # batch of an epoch replays the SAME captured graph as the full ones; other models run it eagerly
        masked = self.use_graph and rem > 0 and self.fused is not None and not self.model.is_lm
        if self.use_graph and (full > 0 or masked):
            self._vision_static(imgs_u8, labels, aug, batch, perms_d, params, epochs, full, max_n, max_epochs,
                                rem if masked else 0)
        for ep in range(epochs):
            for s in range(steps_per_epoch):
                lo = s * batch
                bs = min(batch, n - lo)
                if self.use_graph and (bs == batch or masked):
                    self._replay(("v", batch, tuple(imgs_u8.shape[1:])))
                else:
                    idx = perms_d[ep, lo:lo + bs]
                    x = aug(imgs_u8, idx, params[ep, lo:lo + bs].contiguous())
                    y = labels[idx]
                    self._fwd_bwd({"img": x, "label": y}, float(bs))
                    self._opt_step()
                yield

    def _vision_static(self, imgs_u8, labels, aug, batch, perms_d, params, epochs, full, max_n=0, max_epochs=0,
                       rem=0):
        """(Re)bind the static buffers the captured step reads from, capturing the graph on first use.  ``max_n`` /
        ``max_epochs`` (largest shard / local epochs of the run) size the buffers so they never have to grow.
        ``rem > 0``: every epoch ends with a padded step of ``rem`` real samples (fused ResNet path)."""
        key = ("v", batch, tuple(imgs_u8.shape[1:]))
        n = imgs_u8.shape[0]
        cap = self._static.get(key)
        max_n = max(n, max_n, cap["imgs"].shape[0] if cap else 0)
        max_rows = max(epochs, max_epochs) * ((max_n + batch - 1) // batch)
        if cap is None or cap["imgs"].shape[0] < n or cap["idx"].shape[0] < max_rows:
            cap = {"imgs": torch.zeros((max_n,) + tuple(imgs_u8.shape[1:]), dtype=torch.uint8, device=self.device),
                   "labels": torch.zeros(max_n, dtype=torch.int64, device=self.device),
                   "idx": torch.zeros(max_rows, batch, dtype=torch.int64, device=self.device),
                   "par": torch.zeros(max_rows, batch, 3, dtype=torch.int32, device=self.device),
                   "nv": torch.full((max_rows,), batch, dtype=torch.int32, device=self.device),
                   "row": self.step_ctr}
            self._static[key] = cap
            self.graphs.pop(key, None)
        cap["imgs"][:n].copy_(imgs_u8, non_blocking=True)
        cap["labels"][:n].copy_(labels, non_blocking=True)
        # rows = (epoch, step) pairs in execution order; a padded last step repeats sample 0 of the permutation in
        # its unused slots (real pixels: the padding rows stay finite until the first norm zeroes them)
        spe = full + (1 if rem else 0)
        if rem:
            padn = spe * batch - n
            idx_t = torch.cat([perms_d, perms_d[:, :1].expand(epochs, padn)], dim=1).view(epochs * spe, batch)
            par_t = torch.cat([params, params[:, :1].expand(epochs, padn, 3)], dim=1).view(epochs * spe, batch, 3)
            nv_t = torch.full((epochs, spe), batch, dtype=torch.int32)
            nv_t[:, -1] = rem
        else:
            idx_t = perms_d[:, :full * batch].reshape(epochs * full, batch)
            par_t = params[:, :full * batch].reshape(epochs * full, batch, 3)
            nv_t = torch.full((epochs, spe), batch, dtype=torch.int32)
        rows = epochs * spe
        cap["idx"][:rows].copy_(idx_t, non_blocking=True)
        cap["par"][:rows].copy_(par_t, non_blocking=True)
        cap["nv"][:rows].copy_(nv_t.view(-1), non_blocking=True)
        if key not in self.graphs:
            use_nv = self.fused is not None and not self.model.is_lm

            def step():
                row = cap["row"]
                idx = cap["idx"].index_select(0, row).view(-1)
                par = cap["par"].index_select(0, row).view(-1, 3)
                x = aug(cap["imgs"], idx, par)
                y = cap["labels"].index_select(0, idx)
                nv = cap["nv"].index_select(0, row) if use_nv else None
                self._fwd_bwd({"img": x, "label": y}, float(batch), nv)
                self._opt_step()
                row.add_(1)
            self._capture(key, step)
        return cap

    # -- language model ------------------------------------------------------------------------------------
    def train_lm(self, rows: torch.Tensor, epochs: int, bptt: int):
        """``rows [r, T]`` token matrix of this client; one step per bptt window (reference BatchDataset order)."""
        r, T = rows.shape
        wins = lm_windows(T, bptt)
        n_full = sum(1 for _, l in wins if l == bptt)
        if self.use_graph and n_full > 0:
            key = ("lm", r, bptt)
            cap = self._static.get(key)
            if cap is None or cap["win"].shape[0] < n_full:
                cap = {"win": torch.zeros(n_full, r, bptt, dtype=torch.int64, device=self.device),
                       "row": self.step_ctr}
                self._static[key] = cap
                self.graphs.pop(key, None)
            cap["win"][:n_full].copy_(rows[:, :n_full * bptt].view(r, n_full, bptt).transpose(0, 1),
                                      non_blocking=True)
            if key not in self.graphs:
                def step():
                    tok = cap["win"].index_select(0, cap["row"]).view(r, bptt)
                    self._fwd_bwd({"label": tok}, float(r))
                    self._opt_step()
                    cap["row"].add_(1)
                self._capture(key, step)
        for ep in range(epochs):
            self.step_ctr.zero_()
            if ep:
                self.rng_seed.add_(7919)          # (seed, step) keys the in-kernel dropout / [MASK] draws
            for (s, l) in wins:
                if self.use_graph and l == bptt:
                    self._replay(("lm", r, bptt))
                else:
                    self._fwd_bwd({"label": rows[:, s:s + l].contiguous()}, float(r))
                    self._opt_step()
                    self.step_ctr.add_(1)
                yield

    # -- CUDA graph plumbing -------------------------------------------------------------------------------
    def _capture(self, key, step_fn) -> None:
        """Warm up on a side stream, then capture one full train step.  The warm-up steps really train, so state is
        snapshotted and restored around them."""
        snap = (self.w.clone(), self.wc.clone(), self.mom.clone(), self.meter.buf.clone(), self.step_ctr.clone())
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(3):
                step_fn()
                self.step_ctr.zero_()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        k0 = cu.launch_count()
        if self.fused is not None:
            self._capturing = True
        with torch.cuda.graph(graph, stream=s):
            step_fn()
        if self.fused is not None:
            self._capturing = False
            # keep the captured gradient tensors alive (their addresses are baked into the segment table)
            self._graph_keepalive = getattr(self, "_graph_keepalive", {})
            self._graph_keepalive[key] = self._grads
            self._grads = None
            self._seg_for = getattr(self, "_seg_for", {})
            self._seg_for[key] = torch.tensor(self._captured_segs, dtype=torch.int64)
        self.graph_kernels[key] = cu.launch_count() - k0
        self.graphs[key] = graph
        torch.cuda.synchronize(self.device)
        self.w.copy_(snap[0])
        self.wc.copy_(snap[1])
        self.mom.copy_(snap[2])
        self.meter.buf.copy_(snap[3])
        self.step_ctr.copy_(snap[4])
        self.g.zero_()
        self.scratch.zero_()

    def _replay(self, key) -> None:
        if self.fused is not None and getattr(self, "_seg_active", None) != key:
            # (re)bind the optimiser's segment table to this graph's gradient addresses
            self.seg_tab.copy_(self._seg_for[key])
            self._seg_active = key
        self.graphs[key].replay()
        cu.launch_count_add(self.graph_kernels.get(key, 0))

    # -- results -------------------------------------------------------------------------------------------
    def metrics(self) -> Dict[str, float]:
        return self.meter.result("Local-", lm=self.model.is_lm)
