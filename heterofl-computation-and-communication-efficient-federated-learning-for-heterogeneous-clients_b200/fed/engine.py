"""The federated training engine: one process per GPU, a round's active clients sharded over ranks.

Round = (reference ``src/train_classifier_fed.py:74-94`` / ``train`` :99-124 / ``stats`` :127-138 / ``test`` :141-169)

1. sample active users (``randperm``) and -- in ``dynamic`` mode -- re-draw every user's rate (``multinomial``);
   same RNG calls in the same order as the reference, executed identically on every rank (nothing is broadcast);
2. deal the active clients to ranks (LPT on a per-level cost model) and, per rank, to resident :class:`TrainSlot`
   contexts; ONE distribute launch slices ``W_g`` into all of the rank's starting slots;
3. local training: each slot replays its captured train-step graph on its own stream, slots interleaved;
4. heterogeneous FedAvg: one fused launch; with several ranks the kernel pulls the client arenas of every peer over
   NVLink and publishes the new ``W_g`` shard into every replica (``parallel/peer.py``); the NCCL/gloo path
   (all-reduce of padded sum/count buffers) is kept as the A/B baseline and CPU plumbing;
5. sBN statistics pass + Local/Global test, sharded over ranks;
6. logger / LR schedule / checkpoint in the reference's formats (rank 0).
"""
from __future__ import annotations

import os
import shutil
import time
from collections import deque
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from ..ckpt import compat as ckpt
from ..config.control import model_tag
from ..data.datasets import LMData, fetch_dataset, process_dataset
from ..data.pipeline import VisionAugment
from ..data.split import split_dataset
from ..models import Model
from ..ops import cuda as cu
from ..parallel.dist import DistInfo, all_reduce_sum, barrier, broadcast
from ..parallel.schedule import assign_clients, client_cost
from ..plan.spec import build_spec
from ..utils.logger import Logger
from ..utils.optim import make_scheduler_state
from .evaluator import Evaluator
from .federation import Federation, combine_torch
from .trainer import TrainSlot


@dataclass
class EngineOptions:
    compute_dtype: str = "auto"          # auto: bf16 on CUDA, fp32 on CPU
    use_graph: Optional[bool] = None     # auto: CUDA graphs on GPU
    max_concurrent: int = 16             # resident client slots training concurrently per GPU
    data_on_device: bool = True          # False: dataset stays in pinned host memory, shards are copied every round
    agg_backend: str = "auto"            # auto | p2p | allreduce
    eval_every: int = 1
    out_dir: str = "./output"
    tensorboard: bool = False
    quiet: bool = False
    save: bool = True
    stats_chunk_groups: int = int(os.environ.get("HFL_STATS_CHUNK", "1000"))   # ghost batches per sBN launch
    test_chunk: int = int(os.environ.get("HFL_TEST_CHUNK", "2500"))
    # overlap round r's sBN + test passes with round r+1's training on a low-priority side stream (run_round returns the
    # previous round's metrics as 'prev_eval'; fit() logs them one round late, collect_eval() drains the last one).
    # None = automatic: on from 4 ranks up -- there a GPU holds one or two clients and the evaluation fills SMs the
    # latency-bound training leaves idle (8 GPUs: 231.0 -> 212.7 ms/round); on one GPU the big evaluation kernels only
    # delay the training kernels (no gain measured), so it stays off below 4 ranks.
    async_eval: Optional[bool] = None
    threaded_issue: bool = True          # one host thread per active client slot issues its graph replays
    nvtx: bool = False                   # NVTX range per phase (distribute+train / aggregate / sbn / test) for nsys/ncu
    fault_drop: float = 0.0              # fault injection: each active client is lost with this probability AFTER its
    #                                      local training (its update never reaches the aggregation, the coverage
    #                                      counts shrink accordingly); drawn identically on every rank


class PhaseTimer:
    """CUDA-event (or wall-clock on CPU) phase timing; ``result()`` synchronises once."""

    def __init__(self, device, nvtx: bool = False):
        self.cuda = torch.device(device).type == "cuda"
        self.device = device
        self.marks: List = []
        self.nvtx = bool(nvtx and self.cuda)
        self._open = False

    def open(self, name: str) -> None:
        """Begin an NVTX range for the phase that the next :meth:`mark` closes (no-op unless ``nvtx``)."""
        if self.nvtx:
            if self._open:
                torch.cuda.nvtx.range_pop()
            torch.cuda.nvtx.range_push("hfl/" + name)
            self._open = True

    def mark(self, name: str) -> None:
        if self.nvtx and self._open:
            torch.cuda.nvtx.range_pop()
            self._open = False
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self.marks.append((name, ev))
        else:
            self.marks.append((name, time.perf_counter()))

    def result(self) -> Dict[str, float]:
        out = {}
        if self.cuda and self.marks:
            self.marks[-1][1].synchronize()
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            ms = e0.elapsed_time(e1) if self.cuda else (e1 - e0) * 1e3
            out[n1] = out.get(n1, 0.0) + ms
        return out


class Engine:
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
        if self.is_cuda:
            torch.cuda.manual_seed(self.seed)
        self.dtype = {"auto": torch.bfloat16 if self.is_cuda else torch.float32, "bf16": torch.bfloat16,
                      "fp32": torch.float32}[self.opt.compute_dtype]
        self.use_graph = self.is_cuda if self.opt.use_graph is None else bool(self.opt.use_graph and self.is_cuda)
        # ---- data ----
        self.dataset = dataset if dataset is not None else fetch_dataset(cfg)
        process_dataset(cfg, self.dataset)
        self.is_lm = isinstance(self.dataset["train"], LMData)
        # ---- model ----
        self.spec = build_spec(cfg)
        if not self.spec.federable:
            raise ValueError("Not valid k")  # Bottleneck ResNets cannot be sliced (reference src/fed.py:89)
        self.channels_last = self.is_cuda
        self.host_model = Model(self.spec, cfg, cfg["global_model_rate"], False)
        self.global_model = Model(self.spec, cfg, cfg["global_model_rate"], self.channels_last)
        host_arena = self.host_model.new_arena()                       # consumes the global RNG like a model init
        self.global_arena = self.host_model.convert_arena(host_arena, self.global_model).to(self.device)
        # ---- bookkeeping ----
        self.epoch = 1
        self.logger = Logger(os.path.join(self.opt.out_dir, "runs", "train_%s" % cfg["model_tag"]),
                             tensorboard=self.opt.tensorboard)
        self.sched = make_scheduler_state(cfg, len(self.spec.params))
        self.data_split = None
        self.label_split = None
        self.running: Dict = {}
        self.timings: Dict[str, float] = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._slots: Dict[int, List[TrainSlot]] = {}
        self._models: Dict[float, Model] = {}
        self._peer = None
        self._resumed = False
        self._setup_done = False

    # ------------------------------------------------------------------------------------------------------
    def setup(self, resume_mode: Optional[int] = None) -> None:
        cfg = self.cfg
        mode = cfg.get("resume_mode", 0) if resume_mode is None else resume_mode
        if mode in (1, 2):
            self._resume(mode)
        if self.data_split is None:
            self.data_split, self.label_split = split_dataset(self.dataset, cfg["num_users"], cfg["data_split_mode"],
                                                              cfg.get("classes_size", 0))
        self.federation = Federation(cfg, self.spec, self.global_arena, self.label_split,
                                     channels_last=self.channels_last)
        self.tables = self.federation.tables
        self.evaluator = Evaluator(self.global_model, cfg, self.info, self.dtype, self.opt.stats_chunk_groups,
                                   self.opt.test_chunk)
        self._stage_data()
        self._agg_backend = self._pick_agg_backend()
        self._setup_done = True

    def _pick_agg_backend(self) -> str:
        b = self.opt.agg_backend
        if self.info.world == 1:
            return "local"
        if b == "auto":
            b = "p2p" if (self.is_cuda and self._peer_ok()) else "allreduce"
        if b == "p2p":
            from ..parallel.peer import PeerPool
            need = self._round_arena_bytes()
            self._peer = PeerPool(self.info, need)
            # the global arena moves into the peer pool so that peers can publish W_g shards into it
            ga = self._peer.alloc_persistent(self.global_arena.numel(), torch.float32)
            ga.copy_(self.global_arena)
            self.global_arena = ga
            self.federation.global_arena = ga
            self._peer.enable_all_reduce()
            self._peer.exchange()
            from ..parallel.dist import set_peer_pool
            set_peer_pool(self._peer)       # sBN statistic merge goes over peer memory too (fp32 vectors)
        return b

    def _peer_ok(self) -> bool:
        """P2P aggregation needs every rank on ONE node with peer access between all local GPUs."""
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", self.info.world) or self.info.world)
        if local_world != self.info.world:
            return False
        try:
            n = torch.cuda.device_count()
            me = self.device.index if self.device.index is not None else torch.cuda.current_device()
            return all(torch.cuda.can_device_access_peer(me, j) for j in range(min(n, self.info.world)) if j != me)
        except Exception:
            return False

    def _round_arena_bytes(self) -> int:
        n_act = int(np.ceil(self.cfg["frac"] * self.cfg["num_users"]))
        biggest = self.tables.level_layouts[0].size
        return (n_act * biggest + self.global_arena.numel()) * 4 + (2 << 20)

    # -- data staging --------------------------------------------------------------------------------------
    def _stage_data(self) -> None:
        tr, te = self.dataset["train"], self.dataset["test"]
        dev = self.device
        on_dev = self.opt.data_on_device or not self.is_cuda
        if self.is_lm:
            self.train_rows = tr.token.to(dev) if on_dev else tr.token.pin_memory()
            self.test_rows = te.token.to(dev) if on_dev else te.token.pin_memory()
            return
        pin = (lambda t: t.pin_memory()) if self.is_cuda else (lambda t: t)
        self.train_img = tr.img.to(dev) if on_dev else pin(tr.img)
        self.train_lab = tr.target.to(dev) if on_dev else pin(tr.target)
        self.test_img = te.img.to(dev) if on_dev else pin(te.img)
        self.test_lab = te.target.to(dev) if on_dev else pin(te.target)
        self.aug_train = VisionAugment(tr, True, dev, self.dtype)
        self._max_shard = max(len(v) for v in self.data_split["train"].values())
        U, C = self.cfg["num_users"], self.cfg["classes_size"]
        owner = torch.full((len(te),), -1, dtype=torch.int64)
        for u in range(U):
            owner[torch.as_tensor(self.data_split["test"][u], dtype=torch.long)] = u
        masks = torch.zeros(U, C, dtype=torch.uint8)
        for u in range(U):
            masks[u, torch.as_tensor(list(self.label_split[u]), dtype=torch.long)] = 1
        self.test_owner = owner.to(dev)
        self.user_masks = masks.to(dev)

    def _h2d(self, t: torch.Tensor) -> torch.Tensor:
        if t.device == self.device:
            return t
        self.h2d_bytes += t.numel() * t.element_size()
        return t.to(self.device, non_blocking=True)

    def _src(self, t: torch.Tensor) -> torch.Tensor:
        """Hand a (possibly pinned-host) tensor to the evaluator, which copies it into its static device buffer;
        the H2D bytes are accounted here."""
        if t.device != self.device:
            self.h2d_bytes += t.numel() * t.element_size()
        return t

    # -- slots ---------------------------------------------------------------------------------------------
    def _model_for(self, rate: float) -> Model:
        if rate not in self._models:
            self._models[rate] = Model(self.spec, self.cfg, rate, self.channels_last)
        return self._models[rate]

    def _acquire_slot(self, level: int, busy: set) -> TrainSlot:
        pool = self._slots.setdefault(level, [])
        for s in pool:
            if id(s) not in busy:
                return s
        rate = self.tables.rates[level]
        model = self._model_for(rate)
        if self._use_mega(model):
            from .mega_trainer import MegaSlot
            s = MegaSlot(model, self.cfg, self.device, int(self.cfg["batch_size"]["train"]))
        else:
            s = TrainSlot(model, self.cfg, self.device, self.dtype, self.use_graph)
        pool.append(s)
        return s

    def _use_mega(self, model: Model) -> bool:
        """Resident-kernel training (fed/mega_trainer.py): BatchNorm Block-ResNets in bf16 on a GPU."""
        if not (self.is_cuda and cu.available() and self.dtype == torch.bfloat16 and not self.is_lm):
            return False
        # opt-in (HFL_MEGA=1 or HFL_MEGA_TRAIN=1): measured on one B200 the resident engine trains a round of the
        # headline config in 640 ms against 379 ms for the CUDA-graph path (profiles/mega_engine.md), so the graph path
        # stays the default until the tile pipeline closes that gap
        if os.environ.get("HFL_MEGA_TRAIN", os.environ.get("HFL_MEGA", "0")) != "1" or not self.cfg.get("fused", True):
            return False
        from .mega_trainer import MegaSlot
        return MegaSlot.eligible(model, self.cfg)

    def _mega_grids(self, slots: List) -> None:
        """Split the co-resident CTA budget over the wave's clients: every cooperative grid must fit next to the others
        (a grid that does not fit simply waits for SMs), wide clients get proportionally more CTAs."""
        mega = [s for s in slots if getattr(s, "is_mega", False)]
        if not mega:
            return
        from .mega_trainer import DEFAULT_GRIDS, grid_for_rate
        cap = mega[0].cap
        want = [grid_for_rate(s.model.rate, cap) for s in mega]
        tot = sum(want)
        budget = int(cap * float(os.environ.get("HFL_MEGA_FILL", "0.95")))
        if tot > budget:
            scale = budget / tot
            want = [max(1, int(w * scale)) for w in want]
        elif tot < budget:
            # spare SMs go to the clients by width (more tiles in flight per phase), bounded by what a level can use
            top = {1.0: 148, 0.5: 96, 0.25: 32, 0.125: 12, 0.0625: 6}
            spare = budget - tot
            room = [max(0, top.get(min(top, key=lambda r: abs(r - s.model.rate)), 8) - w) for s, w in zip(mega, want)]
            rt = sum(room)
            if rt > 0:
                want = [w + int(spare * r / rt) if spare < rt else w + r for w, r in zip(want, room)]
        for s, w in zip(mega, want):
            s.grid = int(max(1, min(w, cap)))

    def prewarm(self, slots_per_level: Optional[int] = None) -> int:
        """Create the resident slots and capture their train-step graphs up front (one dummy step each), so no
        capture happens inside a measured round.  Returns the number of slots warmed."""
        if not self._setup_done:
            self.setup()
        if not self.use_graph:
            return 0
        if self._use_mega(self._model_for(self.tables.rates[0])):
            return 0            # resident-kernel training: nothing to capture
        cfg = self.cfg
        n_act = int(np.ceil(cfg["frac"] * cfg["num_users"]))
        per_rank = int(np.ceil(n_act / self.info.world))
        S = min(self.opt.max_concurrent, per_rank) if slots_per_level is None else slots_per_level
        client_levels = sorted({self.tables.level_id(float(r)) for r in cfg["model_rate"]})
        u = 0
        warmed = 0
        for lv in client_levels:
            busy: set = set()
            for _ in range(S):
                s = self._acquire_slot(lv, busy)
                busy.add(id(s))
                s.load_params(torch.zeros_like(s.w).normal_(0, 0.02))
                s.begin_client(0.0, self.federation.user_label_mask(u))
                with torch.cuda.stream(s.stream):
                    g = self._client_generator(s, u, 1, 0)
                    next(g)
                warmed += 1
        torch.cuda.synchronize(self.device)
        return warmed

    def set_data_residency(self, on_device: bool) -> None:
        """Switch between a device-resident dataset and a pinned-host dataset streamed H2D every round."""
        self.opt.data_on_device = bool(on_device)
        self._stage_data()

    # ------------------------------------------------------------------------------------------------------
    def current_lr(self) -> float:
        return float(self.sched.lr)

    def _sync_sampling(self, user_idx: List[int]) -> List[int]:
        """Every rank draws the round's clients and rates from its own default generator (reference RNG order); the
        draws only agree while all ranks consume that generator identically.  Rank 0's draw is authoritative: one
        small broadcast per round removes the dependence on lock-step RNG state (an eager dropout / [MASK] draw on one
        rank would otherwise make ranks silently aggregate different client sets)."""
        fed, info = self.federation, self.info
        rates = np.asarray(fed.model_rate, dtype=np.float64)
        buf = torch.cat([torch.tensor(user_idx, dtype=torch.float64), torch.from_numpy(rates)]).to(self.device)
        buf = broadcast(info, buf, 0).cpu()
        k = len(user_idx)
        fed.model_rate = buf[k:].numpy().astype(np.asarray(fed.model_rate).dtype)
        return [int(u) for u in buf[:k].tolist()]

    def run_round(self, epoch: Optional[int] = None, evaluate: Optional[bool] = None) -> Dict:
        """One communication round; returns metrics + phase timings (ms)."""
        if not self._setup_done:
            self.setup()
        cfg, fed, info = self.cfg, self.federation, self.info
        epoch = self.epoch if epoch is None else epoch
        evaluate = (epoch % max(self.opt.eval_every, 1) == 0) if evaluate is None else evaluate
        timer = PhaseTimer(self.device, self.opt.nvtx)
        timer.mark("start")
        timer.open("train")
        lr = self.current_lr()
        # 1. sampling (same RNG order as the reference: randperm, then multinomial inside distribute)
        user_idx = fed.sample_users()
        fed.make_model_rate()
        if info.world > 1:
            user_idx = self._sync_sampling(user_idx)
        rates = [float(fed.model_rate[u]) for u in user_idx]
        levels = [self.tables.level_id(r) for r in rates]
        M = len(user_idx)
        # 2. schedule
        steps = [self._client_steps(u) for u in user_idx]
        costs = [client_cost(r, s) for r, s in zip(rates, steps)]
        self._assignment = assign_clients(costs, info.world)
        mine = self._assignment[info.rank]
        # 3. local training (per-client RNG streams keyed by (seed, round, user): results do not depend on the
        #    number of ranks or on which rank trains a client)
        results = self._train_clients(user_idx, rates, levels, mine, lr, epoch)
        timer.mark("train")
        # 4. aggregation (minus the clients the fault injector lost)
        timer.open("aggregate")
        lost = self._lost_clients(M, epoch)
        # train metrics (sums over my clients -> all ranks): on the P2P path the merge rides in the aggregation launch
        tm = torch.zeros(4, dtype=torch.float64, device=self.device)
        for r in results.values():
            tm += r["meter"]
        if not self._aggregate(user_idx, levels, mine, results, lost, reduce=tm):
            all_reduce_sum(info, tm)
        timer.mark("aggregate")
        out: Dict = {"epoch": epoch, "lr": lr, "num_active": M, "user_idx": user_idx, "rates": rates,
                     "lost": sorted(lost)}
        # 5. evaluation
        if evaluate and self._async_eval_ok():
            prev = self.collect_eval()        # the previous round's evaluation ran while this round trained
            if prev is not None:
                out["prev_eval"] = prev
            self._launch_eval(epoch)
            out["eval_pending"] = True
            timer.mark("eval_launch")
        elif evaluate:
            shared = torch.Generator().manual_seed(self.seed * 1000003 + epoch * 7919 + 104729)
            if self.is_lm:
                timer.open("test")
                test = self.evaluator.test_lm(self.global_arena, self._h2d(self.test_rows))
                timer.mark("test")
            else:
                timer.open("sbn")
                self.running = self.evaluator.sbn_stats(self.global_arena, self.dataset["train"],
                                                        self._src(self.train_img), shared)
                timer.mark("sbn")
                timer.open("test")
                test = self.evaluator.test_vision(self.global_arena, self.running, self.dataset["test"],
                                                  self._src(self.test_img), self._src(self.test_lab),
                                                  self.test_owner, self.user_masks)
                timer.mark("test")
            out["test"] = {k: v for k, v in test.items() if not k.startswith("_")}
            out["_test_n"] = {k: v for k, v in test.items() if k.startswith("_")}
        t = tm.tolist()                                       # device -> host read of the round's result
        self.d2h_bytes += 32
        if getattr(self, "_mega_launched", False):
            c = self._mega_ctl.tolist()
            if c[1] != 0:
                raise RuntimeError("training megakernel aborted (code %d, diag %s)" % (c[1], c[8:14]))
        if self._peer is not None:
            self._peer.check()                                # a peer that never reached a device barrier -> error
        n = max(t[2], 1.0)
        train = {"Local-Loss": t[0] / n}
        if self.is_lm:
            train["Local-Perplexity"] = t[3] / n
        else:
            train["Local-Accuracy"] = 100.0 * t[1] / n
        out["train"], out["train_n"] = train, t[2]
        out["timings_ms"] = timer.result()
        return out

    # -- asynchronous evaluation ----------------------------------------------------------------------------
    def _async_eval_ok(self) -> bool:
        want = self.opt.async_eval
        if want is None:
            # vision models only: the fused masked-LM evaluation forward on the helper thread, concurrent with the next
            # round's training, ended in a peer-barrier timeout on 2 and 4 ranks (gpurun_out/r2m, r2o; root cause not
            # found -- with the functional evaluation the asynchronous mode works: 403 ms vs 418 ms synchronous on 2
            # ranks).  test_lm therefore uses the fused forward only in the synchronous mode, and the automatic mode
            # leaves the LM evaluation inside its round.
            want = self.info.world >= 4 and not self.is_lm
        return bool(want and self.is_cuda and self._agg_backend in ("local", "p2p"))

    def _launch_eval(self, epoch: int) -> None:
        """Snapshot ``W_g`` and run the sBN + test passes on a low-priority side stream, enqueued from a helper host
        thread (graph launches block the issuing thread once the stream's queue is full); the next round's training
        proceeds on the main thread / slot streams meanwhile.  Results are fetched by :meth:`collect_eval`."""
        import threading
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_eval_stream", None) is None:
            self._eval_stream = torch.cuda.Stream(device=self.device, priority=0)
            self._eval_arena = torch.empty_like(self.global_arena)
            self._eval_done = None
            self._eval_result = None
            self._eval_thread = None
        self._join_eval_thread()
        if self._eval_done is not None:
            main.wait_event(self._eval_done)           # previous evaluation must have consumed the snapshot
        self._eval_arena.copy_(self.global_arena, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        self._eval_error = None
        self._eval_thread = threading.Thread(target=self._eval_worker, args=(epoch, ready), daemon=True)
        self._eval_thread.start()

    def _eval_worker(self, epoch: int, ready) -> None:
        from ..parallel.dist import peer_channel
        try:
            torch.cuda.set_device(self.device)
            shared = torch.Generator().manual_seed(self.seed * 1000003 + epoch * 7919 + 104729)
            with torch.cuda.stream(self._eval_stream), peer_channel(1), torch.no_grad():
                self._eval_stream.wait_event(ready)
                if self.is_lm:
                    test = self.evaluator.test_lm(self._eval_arena, self._h2d(self.test_rows), sync=False)
                else:
                    self.running = self.evaluator.sbn_stats(self._eval_arena, self.dataset["train"],
                                                            self._src(self.train_img), shared)
                    test = self.evaluator.test_vision(self._eval_arena, self.running, self.dataset["test"],
                                                      self._src(self.test_img), self._src(self.test_lab),
                                                      self.test_owner, self.user_masks, sync=False)
                done = torch.cuda.Event()
                done.record(self._eval_stream)
            self._eval_done = done
            self._eval_result = (epoch, test)
        except BaseException as e:      # surfaced by collect_eval on the main thread
            self._eval_error = e

    def _join_eval_thread(self) -> None:
        t = getattr(self, "_eval_thread", None)
        if t is not None:
            t.join()
            self._eval_thread = None
            if self._eval_error is not None:
                raise self._eval_error

    def collect_eval(self) -> Optional[Dict]:
        """Wait for the most recently launched evaluation and return ``{'epoch', 'test': {...}}`` (None if none)."""
        self._join_eval_thread()
        if getattr(self, "_eval_result", None) is None:
            return None
        self._eval_done.synchronize()
        epoch, pending = self._eval_result
        self._eval_result = None
        test = self.evaluator.finish(pending)
        self.d2h_bytes += 48
        return {"epoch": epoch, "test": {k: v for k, v in test.items() if not k.startswith("_")},
                "_test_n": {k: v for k, v in test.items() if k.startswith("_")}}

    def _client_steps(self, u: int) -> int:
        cfg = self.cfg
        if self.is_lm:
            T = self.dataset["train"].token.shape[1]
            return cfg["num_epochs"]["local"] * int(np.ceil(T / cfg["bptt"]))
        n = len(self.data_split["train"][u])
        return cfg["num_epochs"]["local"] * int(np.ceil(n / cfg["batch_size"]["train"]))

    # -- local training ------------------------------------------------------------------------------------
    def _train_clients(self, user_idx, rates, levels, mine, lr, epoch) -> Dict[int, Dict]:
        cfg, fed = self.cfg, self.federation
        if mine and all(self._use_mega(self._model_for(rates[i])) for i in mine):
            return self._train_clients_mega(user_idx, rates, levels, mine, lr, epoch)
        main = torch.cuda.current_stream(self.device) if self.is_cuda else None
        order = sorted(mine, key=lambda i: -client_cost(rates[i], self._client_steps(user_idx[i])))
        pending = deque(order)
        active: List = []
        busy: set = set()
        results: Dict[int, Dict] = {}
        result_arenas = self._result_arenas([levels[i] for i in mine], mine)
        epochs = cfg["num_epochs"]["local"]

        def start_batch(batch_pos: List[int]) -> None:
            slots = []
            for i in batch_pos:
                s = self._acquire_slot(levels[i], busy)
                busy.add(id(s))
                slots.append(s)
            # ONE distribute launch for the whole batch of starting clients
            if self.is_cuda and cu.available():
                from ..ops import fed_ops
                fed_ops.distribute(self.tables, self.global_arena, [s.w for s in slots], [levels[i] for i in batch_pos],
                                   [s.wc for s in slots] if slots[0].mixed else None, [s.mom for s in slots])
                for s in slots:
                    s.g.zero_()
            else:
                arenas = fed.distribute([user_idx[i] for i in batch_pos], redraw=False)
                for s, a in zip(slots, arenas):
                    s.load_params(a)
            self._mega_grids(slots)
            for i, s in zip(batch_pos, slots):
                u = user_idx[i]
                s.begin_client(lr, fed.user_label_mask(u), rng_seed=(self.seed * 1000003 + epoch * 7919) * 4099 + int(u))
                g = self._client_generator(s, u, epochs, epoch)
                if self.is_cuda:
                    s.stream.wait_stream(main)
                active.append([g, s, i])

        if self.is_cuda and self.opt.threaded_issue:
            # One host thread per active slot drives that slot's graph replays.  cudaGraphLaunch of graphs whose
            # kernel nodes carry TMA descriptors (cuDNN's sm_100 kernels, our tcgen05 kernels) blocks its caller for
            # hundreds of microseconds; issued round-robin from ONE thread those stalls serialise all slots.
            import threading
            errors: List[BaseException] = []

            def worker(g, s, i):
                try:
                    torch.cuda.set_device(self.device)
                    with torch.cuda.stream(s.stream):
                        for _ in g:
                            pass
                        result_arenas[i].copy_(s.w, non_blocking=True)
                        meter = s.meter.buf.clone()
                    results[i] = {"arena": result_arenas[i], "meter": meter, "level": levels[i]}
                except BaseException as e:       # re-raised on the main thread
                    errors.append(e)

            while pending:
                wave = [pending.popleft() for _ in range(min(self.opt.max_concurrent, len(pending)))]
                start_batch(wave)
                threads = []
                for g, s, i in active:
                    with torch.cuda.stream(s.stream):
                        next(g, None)            # first step on the main thread: static-buffer setup / graph capture
                    t = threading.Thread(target=worker, args=(g, s, i), daemon=True)
                    t.start()
                    threads.append(t)
                for t in threads:
                    t.join()
                for g, s, i in active:
                    main.wait_stream(s.stream)
                    busy.discard(id(s))
                active.clear()
                if errors:
                    raise errors[0]
            return results
        while pending or active:
            room = self.opt.max_concurrent - len(active)
            if pending and room > 0:
                start_batch([pending.popleft() for _ in range(min(room, len(pending)))])
            for item in list(active):
                g, s, i = item
                try:
                    if self.is_cuda:
                        with torch.cuda.stream(s.stream):
                            next(g)
                    else:
                        next(g)
                except StopIteration:
                    if self.is_cuda:
                        with torch.cuda.stream(s.stream):
                            result_arenas[i].copy_(s.w, non_blocking=True)
                            meter = s.meter.buf.clone()
                        main.wait_stream(s.stream)
                    else:
                        result_arenas[i].copy_(s.w)
                        meter = s.meter.buf.clone()
                    results[i] = {"arena": result_arenas[i], "meter": meter, "level": levels[i]}
                    busy.discard(id(s))
                    active.remove(item)
        return results

    def _client_generator(self, s: TrainSlot, u: int, epochs: int, epoch: int):
        cfg = self.cfg
        if self.is_lm:
            ridx = torch.as_tensor(self.data_split["train"][u])
            if self.train_rows.device == self.device:
                rows = self.train_rows[ridx.to(self.device)]
            else:
                rows = self._h2d(self.train_rows[ridx].pin_memory() if self.is_cuda else self.train_rows[ridx])
            return s.train_lm(rows, epochs, int(cfg["bptt"]))
        imgs, labs, gen = self._client_data(u, epoch)
        return s.train_vision(imgs, labs, self.aug_train, epochs, int(cfg["batch_size"]["train"]), gen,
                              max_n=self._max_shard, max_epochs=int(cfg["num_epochs"]["local"]))

    def _client_data(self, u: int, epoch: int):
        """(images uint8 [n,H,W,C], labels, per-client RNG) of user ``u`` on the device (H2D bytes accounted)."""
        idx = torch.as_tensor(self.data_split["train"][u], dtype=torch.long)
        if self.train_img.device == self.device:
            di = idx.to(self.device)
            imgs, labs = self.train_img[di], self.train_lab[di]
        else:
            imgs = self._h2d(self.train_img[idx].pin_memory() if self.is_cuda else self.train_img[idx])
            labs = self._h2d(self.train_lab[idx].pin_memory() if self.is_cuda else self.train_lab[idx])
        gen = torch.Generator().manual_seed((self.seed * 1000003 + epoch * 7919) * 4099 + int(u))
        return imgs, labs, gen

    def _train_clients_mega(self, user_idx, rates, levels, mine, lr, epoch) -> Dict[int, Dict]:
        """Resident-kernel path: the clients of a wave (up to 16) train inside ONE cooperative launch -- lock-step
        phase programs, shared grid barriers, tiles of all clients interleaved over all SMs (csrc/mega_kernel.cu)."""
        from .mega_trainer import launch_wave
        from ..ops import fed_ops
        cfg, fed = self.cfg, self.federation
        results: Dict[int, Dict] = {}
        result_arenas = self._result_arenas([levels[i] for i in mine], mine)
        order = sorted(mine, key=lambda i: -client_cost(rates[i], self._client_steps(user_idx[i])))
        epochs, batch = int(cfg["num_epochs"]["local"]), int(cfg["batch_size"]["train"])
        K = min(int(cu.C().mega_max_jobs()), max(int(self.opt.max_concurrent), 1))
        if getattr(self, "_mega_ctl", None) is None:
            self._mega_ctl = torch.zeros(16, dtype=torch.int32, device=self.device)
            self._mega_cap = int(cu.C().mega_max_grid())
        for w0 in range(0, len(order), K):
            wave = order[w0:w0 + K]
            busy: set = set()
            slots = []
            for i in wave:
                s = self._acquire_slot(levels[i], busy)
                busy.add(id(s))
                slots.append(s)
            fed_ops.distribute(self.tables, self.global_arena, [s.w for s in slots], [levels[i] for i in wave],
                               [s.wc for s in slots], [s.mom for s in slots])
            for i, s in zip(wave, slots):
                u = user_idx[i]
                s.g.zero_()
                s.begin_client(lr, fed.user_label_mask(u))
                imgs, labs, gen = self._client_data(u, epoch)
                s.prepare_client(imgs, labs, self.aug_train, epochs, batch, gen, max_n=self._max_shard,
                                 max_epochs=epochs)
            launch_wave(slots, int(os.environ.get("HFL_MEGA_GRID", "0")) or self._mega_cap, self._mega_ctl)
            self._mega_launched = True
            for i, s in zip(wave, slots):
                result_arenas[i].copy_(s.w, non_blocking=True)
                results[i] = {"arena": result_arenas[i], "meter": s.meter.buf.clone(), "level": levels[i]}
        return results

    def _result_arenas(self, my_levels: List[int], mine: List[int]) -> Dict[int, torch.Tensor]:
        out = {}
        if self._peer is not None:
            self._peer.reset_round()
        for i, lv in zip(mine, my_levels):
            n = self.tables.level_layouts[lv].size
            if self._peer is not None:
                out[i] = self._peer.alloc_round(n, torch.float32)
            else:
                out[i] = torch.empty(n, dtype=torch.float32, device=self.device)
        return out

    # -- aggregation ---------------------------------------------------------------------------------------
    def _lost_clients(self, M: int, epoch: int) -> set:
        p = float(self.opt.fault_drop)
        if p <= 0.0:
            return set()
        g = torch.Generator().manual_seed(self.seed * 1000003 + epoch * 7919 + 15485863)
        lost = set(torch.nonzero(torch.rand(M, generator=g) < p).flatten().tolist())
        if len(lost) == M:                                    # keep one survivor so the round still moves W_g
            lost.discard(min(lost))
        return lost

    def _aggregate(self, user_idx, levels, mine, results, lost=(), reduce=None) -> bool:
        """Returns True when ``reduce`` (the round's metric vector) was summed over the ranks by the aggregation itself."""
        fed, info = self.federation, self.info
        masks_all = [fed.user_label_mask(u) for u in user_idx]
        if lost:
            results = {i: r for i, r in results.items() if i not in lost}
        if self._agg_backend == "local":
            order = sorted(results.keys())
            arenas = [results[i]["arena"] for i in order]
            if self.is_cuda and cu.available():
                from ..ops import fed_ops
                fed_ops.aggregate(self.tables, self.global_arena, arenas, [levels[i] for i in order],
                                  [masks_all[i] for i in order])
            else:
                combine_torch(self.tables, self.global_arena, arenas, [self.tables.rates[levels[i]] for i in order],
                              [masks_all[i] for i in order])
        elif self._agg_backend == "p2p":
            from ..parallel.peer import p2p_aggregate
            p2p_aggregate(self._peer, self.tables, self.global_arena, levels, self._assignment, masks_all, info,
                          active=[i for i in range(len(levels)) if i not in lost] if lost else None, reduce=reduce)
            return reduce is not None
        else:
            from ..parallel.baseline import allreduce_aggregate
            allreduce_aggregate(self.tables, self.global_arena, levels, results, masks_all, info)
        return False

    # -- full training loop --------------------------------------------------------------------------------
    def fit(self, num_rounds: Optional[int] = None) -> None:
        if not self._setup_done:
            self.setup()
        cfg = self.cfg
        last = cfg["num_epochs"]["global"] if num_rounds is None else min(cfg["num_epochs"]["global"],
                                                                          self.epoch + num_rounds - 1)
        crash_at = int(os.environ.get("HFL_FAULT_CRASH_ROUND", "0") or 0)   # fault injection for the supervisor test
        while self.epoch <= last:
            ep = self.epoch
            if crash_at and ep == crash_at and not self._resumed:
                os._exit(17)                                   # simulate a rank dying mid-run (no cleanup, like SIGKILL)
            self.logger.safe(True)
            t0 = time.time()
            res = self.run_round(ep)
            dt = time.time() - t0
            if "prev_eval" in res:             # asynchronous evaluation: round ep-1's test metrics arrive now
                self._log_test(res["prev_eval"])
            self._log_round(res, dt)
            # ReduceLROnPlateau needs the monitored value (reference passes logger.mean['train/<pivot>'])
            pm = str(cfg.get("pivot_metric", ""))
            mean = getattr(self.logger, "mean", {})
            metric = mean.get("train/" + pm, mean.get("test/" + pm, mean.get("train/Local-Loss")))
            self.sched.step(metric=metric)
            self.logger.safe(False)
            if self.opt.save and self.info.is_main:
                self.save_checkpoint(ep + 1, res)
            self.logger.reset()
            self.epoch = ep + 1
            barrier(self.info)
        tail = self.collect_eval() if self._async_eval_ok() else None     # last round's asynchronous evaluation
        if tail is not None:
            self.logger.safe(True)
            self._log_test(tail)
            self.logger.safe(False)
            if self.opt.save and self.info.is_main:
                self.save_checkpoint(self.epoch, None)
            self.logger.reset()
        self.logger.safe(False)

    def _log_round(self, res: Dict, dt: float) -> None:
        cfg, lg = self.cfg, self.logger
        names = cfg["metric_name"]
        lg.append(res["train"], "train", n=max(res["train_n"], 1.0))
        left = (cfg["num_epochs"]["global"] - res["epoch"]) * dt
        info = {"info": ["Model: {}".format(cfg["model_tag"]),
                         "Train Epoch: {}(100%)".format(res["epoch"]),
                         "Clients: {}".format(res["num_active"]),
                         "Learning rate: {}".format(res["lr"]),
                         "Round time: {:.3f}s".format(dt),
                         "Experiment Finished Time: {:.0f}s".format(left)]}
        tms = res.get("timings_ms") or {}
        if tms:     # observability beyond the reference's wall-clock ETA: per-phase device time and throughput
            info["info"].append("Phases(ms): " + " ".join("{}={:.1f}".format(k, v) for k, v in tms.items()))
            if dt > 0:
                info["info"].append("Samples/s: {:.0f}".format(res["train_n"] / dt))
        lg.append(info, "train", mean=False)
        quiet = self.opt.quiet or not self.info.is_main
        lg.write("train", names["train"]["Local"], quiet=quiet)
        if "test" in res:
            self._log_test(res)

    def _log_test(self, res: Dict) -> None:
        """Log one round's evaluation (``res`` = a run_round result or a collect_eval / 'prev_eval' record)."""
        cfg, lg = self.cfg, self.logger
        names = cfg["metric_name"]
        quiet = self.opt.quiet or not self.info.is_main
        tn = res.get("_test_n", {})
        loc = {k: v for k, v in res["test"].items() if k.startswith("Local-")}
        glo = {k: v for k, v in res["test"].items() if k.startswith("Global-")}
        if loc:
            lg.append(loc, "test", n=max(tn.get("_n_local", 1.0), 1.0))
        lg.append(glo, "test", n=max(tn.get("_n_global", 1.0), 1.0))
        lg.append({"info": ["Model: {}".format(cfg["model_tag"]),
                            "Test Epoch: {}(100%)".format(res["epoch"])]}, "test", mean=False)
        tnames = names["test"].get("Local", []) + names["test"]["Global"]
        lg.write("test", tnames, quiet=quiet)

    # -- checkpointing (reference format) ------------------------------------------------------
    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return self.global_model.layout.state_dict(self.global_arena.detach().float().cpu())

    def load_state_dict(self, sd, strict: bool = True) -> None:
        host = torch.zeros(self.global_model.layout.size)
        cur = self.global_arena.detach().float().cpu()
        host.copy_(cur)
        self.global_model.layout.load_state_dict(host, sd, strict=strict)
        self.global_arena.copy_(host.to(self.device))

    def save_checkpoint(self, next_epoch: int, res: Optional[Dict] = None) -> str:
        cfg = self.cfg
        path = os.path.join(self.opt.out_dir, "model", "%s_checkpoint.pt" % cfg["model_tag"])
        save_cfg = {k: v for k, v in cfg.items()}
        obj = {"cfg": save_cfg, "epoch": next_epoch, "data_split": self.data_split, "label_split": self.label_split,
               "model_dict": self.state_dict(), "optimizer_dict": self.sched.optimizer.state_dict(),
               "scheduler_dict": self.sched.scheduler.state_dict(), "logger": self.logger}
        ckpt.save(obj, path)
        pm = "test/%s" % cfg["pivot_metric"]
        if pm in self.logger.mean:
            cur = self.logger.mean[pm]
            better = cur > cfg["pivot"] if cfg["pivot_metric"].endswith("Accuracy") else cur < cfg["pivot"]
            if better:
                cfg["pivot"] = cur
                ckpt.atomic_copy(path, os.path.join(self.opt.out_dir, "model", "%s_best.pt" % cfg["model_tag"]))
        return path

    def _resume(self, mode: int, load_tag: str = "checkpoint") -> None:
        cfg = self.cfg
        path = os.path.join(self.opt.out_dir, "model", "%s_%s.pt" % (cfg["model_tag"], load_tag))
        if not os.path.exists(path):
            if self.info.is_main and not self.opt.quiet:
                print("Not exists model tag: {}, start from scratch".format(cfg["model_tag"]))
            return
        ck = None
        for cand in (path, os.path.join(self.opt.out_dir, "model", "%s_best.pt" % cfg["model_tag"])):
            try:
                ck = ckpt.load(cand)
                break
            except Exception as e:      # truncated / corrupt file (e.g. a crash before the atomic-rename fix)
                if self.info.is_main:
                    print("Could not load checkpoint {}: {}".format(cand, e))
        if ck is None:
            if self.info.is_main:
                print("No loadable checkpoint for {}, start from scratch".format(cfg["model_tag"]))
            return
        self.data_split, self.label_split = ck["data_split"], ck["label_split"]
        self.load_state_dict(ck["model_dict"])
        if mode == 1:
            self.epoch = ck["epoch"]
            self._resumed = True
            self.sched.optimizer.load_state_dict(ck["optimizer_dict"])
            self.sched.scheduler.load_state_dict(ck["scheduler_dict"])
            self.logger = ck["logger"]
            if self.info.is_main and not self.opt.quiet:
                print("Resume from {}".format(self.epoch))
        else:
            self.logger = Logger(os.path.join(self.opt.out_dir, "runs", "%s" % cfg["model_tag"]),
                                 tensorboard=self.opt.tensorboard)
