"""Reference arm of ``bench.py``: drives the UNMODIFIED reference sources in ``baseline/_ref/src`` through their own
entry-point functions (``train_classifier_fed.train / stats / test`` resp. the transformer equivalents) -- none of this
framework's models, kernels or engine are on that path.

What this file adds around the reference (it cannot run standalone on this box):
* synthetic CIFAR-10 / MNIST / WikiText-2 shaped data written in the reference's own *processed-cache* format
  (``./data/<NAME>/processed/{train,test,meta}.pt``) so its stock dataset classes, PIL transforms and DataLoaders run
  unchanged (there is no network to download the real files);
* a stub ``anytree`` module (imported by ``datasets/*.py`` but only used while *building* the cache);
* for torch >= 2, a shim for ``nn.TransformerEncoder.forward`` probing ``layers[0].self_attn`` (LM config only);
* for N > 1 GPUs (the reference has no distributed mode): the thin harness BASELINE.md describes -- active clients are
  dealt round-robin to ranks, every rank runs the reference's ``Local.train`` for its clients, the trained local
  tensors are broadcast from their owner with NCCL, and every rank then calls the reference's own
  ``Federation.combine``; the sBN/test phases stay on every rank exactly as in the reference (they are not
  parallelised there).
"""
from __future__ import annotations

import copy
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")


def _fabricate_data(data_name: str, seed: int = 0) -> None:
    """Write synthetic processed caches the reference dataset classes load as-is."""
    import torch
    sys.path.insert(0, ROOT)
    from hfl_b200.data.datasets import synth_lm, synth_vision
    proc = os.path.join(REF_SRC, "data", data_name, "processed")
    if os.path.exists(os.path.join(proc, "meta.pt")):
        return
    os.makedirs(proc, exist_ok=True)
    if data_name in ("CIFAR10", "MNIST"):
        for split in ("train", "test"):
            ds = synth_vision(data_name, split, None, seed)
            img = ds.img.numpy()
            if data_name == "MNIST":
                img = img[..., 0]
            torch.save((img, {"label": ds.target.numpy()}), os.path.join(proc, "%s.pt" % split), pickle_protocol=2)
        torch.save(({"label": None}, {"label": 10}), os.path.join(proc, "meta.pt"), pickle_protocol=2)
    else:
        # Vocab must unpickle as the reference's own datasets.lm.Vocab: build it with the reference class
        import datasets.lm as rlm  # noqa (reference module; sys.path already points at REF_SRC)
        v = rlm.Vocab()
        tr = synth_lm(data_name, "train", None, None, seed)
        for i in range(2, len(tr.vocab)):
            v.add("tok%d" % i)
        for split in ("train", "valid", "test"):
            ds = tr if split == "train" else synth_lm(data_name, "test", None, None, seed)
            torch.save(ds.token, os.path.join(proc, "%s.pt" % split), pickle_protocol=2)
        torch.save(v, os.path.join(proc, "meta.pt"), pickle_protocol=2)


def run(args, data_name: str, model_name: str, control: str, ClockSampler) -> None:
    # the reference calls torch.load() on numpy/object pickles without weights_only=False (torch < 2.6 semantics)
    os.environ["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if "anytree" not in sys.modules:
        try:
            import anytree  # noqa: F401
        except Exception:
            sys.modules["anytree"] = types.ModuleType("anytree")
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    is_lm = model_name == "transformer"
    entry = "train_transformer_fed" if is_lm else "train_classifier_fed"
    sys.argv = [entry + ".py", "--data_name", data_name, "--model_name", model_name, "--control_name", control,
                "--device", "cuda" if use_cuda else "cpu"]
    if rank == 0:
        _fabricate_data(data_name)
    if world > 1:
        dist.barrier()
    # the repo root carries same-named wrapper scripts (train_classifier_fed.py, ...): make sure the reference's own
    # modules win every `import` from here on
    sys.path[:] = [REF_SRC] + [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, REF_SRC)]
    for name in (entry, "config", "data", "fed", "models", "utils", "logger", "metrics", "modules", "datasets"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REF_SRC):
            del sys.modules[name]
    T = __import__(entry)                 # runs the reference's argparse/cfg bootstrap at import
    assert str(T.__file__).startswith(REF_SRC), T.__file__
    cfg = T.cfg
    T.process_control()
    cfg["model_tag"] = "0_%s_label_%s_%s" % (data_name, model_name, control)
    if args.local_epochs:
        cfg["num_epochs"]["local"] = args.local_epochs
    if is_lm:
        # torch>=2: nn.TransformerEncoder.forward probes attributes of stock layers; run the reference's own layers
        # in sequence instead (same modules, same parameter names)
        import torch.nn as nn

        def _enc_forward(self, src, mask=None, src_key_padding_mask=None, is_causal=None):
            out = src
            for mod in self.layers:
                out = mod(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask)
            return out
        nn.TransformerEncoder.forward = _enc_forward
    torch.manual_seed(0)
    if use_cuda:
        torch.cuda.manual_seed(0)
    dataset = T.fetch_dataset(cfg["data_name"], cfg["subset"])
    T.process_dataset(dataset)
    model = eval('T.models.{}(model_rate=cfg["global_model_rate"]).to(cfg["device"])'.format(cfg["model_name"]))
    optimizer = T.make_optimizer(model, cfg["lr"])
    scheduler = T.make_scheduler(optimizer)
    data_split, label_split = T.split_dataset(dataset, cfg["num_users"], cfg["data_split_mode"])
    logger = T.Logger(os.path.join("/tmp", "hfl_ref_runs", "train_%s_%d" % (cfg["model_tag"], rank)))
    federation = T.Federation(model.state_dict(), cfg["model_rate"], label_split)
    samples_box = [0.0]

    import contextlib
    autocast = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if (getattr(args, "ref_autocast", False) and use_cuda) \
        else contextlib.nullcontext
    shard_eval = bool(getattr(args, "ref_shard_eval", False)) and world > 1 and not is_lm

    def one_round(epoch: int) -> None:
        with autocast():
            _one_round(epoch)

    def _one_round(epoch: int) -> None:
        logger.safe(True)
        if world == 1:
            T.train(dataset["train"], data_split["train"], label_split, federation, model, optimizer, logger, epoch)
        else:
            _train_sharded(T, dist, rank, world, dataset["train"], data_split["train"], label_split, federation,
                           model, optimizer, logger)
        samples_box[0] += logger.counter.get("train/Local-Loss", 0)
        if not args.no_eval:
            if is_lm:
                T.test(dataset["test"], model, logger, epoch)
            elif shard_eval:
                _eval_sharded(T, dist, rank, world, dataset, data_split["test"], label_split, model, logger)
            else:
                test_model = T.stats(dataset["train"], model)
                T.test(dataset["test"], data_split["test"], label_split, test_model, logger, epoch)
        scheduler.step()
        logger.safe(False)
        logger.reset()

    ep = 1
    devnull = open(os.devnull, "w")
    stdout = sys.stdout
    sys.stdout = devnull                  # the reference prints progress lines; keep stdout to ONE JSON line
    try:
        for _ in range(args.warmup):
            one_round(ep)
            ep += 1
        if world > 1:
            dist.barrier()
        if use_cuda:
            torch.cuda.synchronize()
        samples_box[0] = 0.0
        if use_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as cs:
            t0 = time.perf_counter()
            if use_cuda:
                e0.record()
            for _ in range(args.steps):
                one_round(ep)
                ep += 1
            if use_cuda:
                e1.record()
                torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            dist.barrier()
    finally:
        sys.stdout = stdout
    dev = "cuda" if use_cuda else "cpu"
    ms = torch.tensor([e0.elapsed_time(e1) if use_cuda else wall_ms], device=dev, dtype=torch.float64)
    smp = torch.tensor([samples_box[0]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(smp, op=dist.ReduceOp.SUM)
    ms, samples = float(ms.item()), float(smp.item())
    if rank == 0:
        val = samples / (ms / 1e3)
        line = {"impl": "reference", "metric": "federated_train_samples_per_sec", "value": val, "unit": "samples/s",
                "rounds_per_sec": args.steps / (ms / 1e3), "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16-autocast" if getattr(args, "ref_autocast", False) else "fp32", "data": "synthetic",
                "eval_sharded": shard_eval,
                "config": {"model": "%s %s %s" % (data_name, model_name, control),
                           "global_batch": cfg["batch_size"]["train"], "seq_len": None,
                           "parallelism": "reference serial client loop" if world == 1 else
                           "reference Local.train sharded over %d ranks + NCCL broadcast + reference combine" % world,
                           "round": "distribute+local-train+aggregate" + ("" if args.no_eval else "+sBN-stats+test"),
                           "train_samples_per_round": samples / args.steps,
                           "local_epochs": cfg["num_epochs"]["local"]},
                "clocks": cs.summary(),
                "e2e": {"value": val, "unit": "samples/s", "note": "the reference pipeline is end-to-end by "
                        "construction (CPU DataLoader -> H2D per step, .item() per step)"},
                "gpu_launches": 0}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _train_sharded(T, dist, rank, world, dataset, data_split, label_split, federation, global_model, optimizer,
                   logger) -> None:
    """Reference ``train()`` with the client loop sharded over ranks (the harness BASELINE.md describes)."""
    import torch
    cfg = T.cfg
    global_model.load_state_dict(federation.global_parameters)
    global_model.train(True)
    # identical RNG stream on every rank up to here (same seed) => same user_idx / rates everywhere
    local, local_parameters, user_idx, param_idx = T.make_local(dataset, data_split, label_split, federation)
    lr = optimizer.param_groups[0]["lr"]
    rng_state = torch.get_rng_state()
    for m in range(len(local)):
        if m % world == rank:
            local_parameters[m] = copy.deepcopy(local[m].train(local_parameters[m], lr, logger))
    torch.set_rng_state(rng_state)        # keep the ranks' sampling streams aligned for the next round
    for m in range(len(local)):
        for k in local_parameters[m]:
            t = local_parameters[m][k]
            if not t.is_cuda:
                t = t.cuda()
            t = t.contiguous()
            dist.broadcast(t, src=m % world)
            local_parameters[m][k] = t
    federation.combine(local_parameters, param_idx, user_idx)
    global_model.load_state_dict(federation.global_parameters)


def _eval_sharded(T, dist, rank, world, dataset, test_split, label_split, model, logger) -> None:
    """``--ref-shard-eval``: the reference's sBN pass and test loops with the data dealt to the ranks (the stock
    reference runs both on every rank; this variant exists so that the multi-GPU ratio is not inflated by an
    un-parallelised evaluation).  Uses the reference's own model classes, DataLoaders, collate and Metric."""
    import torch
    cfg = T.cfg
    with torch.no_grad():
        # ---- sBN: every rank accumulates BN running statistics over its 1/world of the train set, then the
        # cumulative averages (momentum=None) are averaged over the ranks (equal shard sizes)
        test_model = eval('T.models.{}(model_rate=cfg["global_model_rate"], track=True).to(cfg["device"])'
                          .format(cfg["model_name"]))
        test_model.load_state_dict(model.state_dict(), strict=False)
        n = len(dataset["train"])
        idx = list(range(rank, n, world))
        loader = T.make_data_loader({"train": T.SplitDataset(dataset["train"], idx)})["train"]
        test_model.train(True)
        for inp in loader:
            inp = T.to_device(T.collate(inp), cfg["device"])
            test_model(inp)
        for name, buf in test_model.named_buffers():
            if buf.dtype.is_floating_point:
                dist.all_reduce(buf)
                buf.div_(world)
        # ---- test: users dealt round-robin for the Local metrics, the global test set in strided slices
        metric = T.Metric()
        test_model.train(False)
        for m in range(rank, cfg["num_users"], world):
            loader = T.make_data_loader({"test": T.SplitDataset(dataset["test"], test_split[m])})["test"]
            for inp in loader:
                inp = T.collate(inp)
                bs = inp["img"].size(0)
                inp["label_split"] = torch.tensor(label_split[m])
                inp = T.to_device(inp, cfg["device"])
                out = test_model(inp)
                logger.append(metric.evaluate(cfg["metric_name"]["test"]["Local"], inp, out), "test", bs)
        gidx = list(range(rank, len(dataset["test"]), world))
        loader = T.make_data_loader({"test": T.SplitDataset(dataset["test"], gidx)})["test"]
        for inp in loader:
            inp = T.collate(inp)
            bs = inp["img"].size(0)
            inp = T.to_device(inp, cfg["device"])
            out = test_model(inp)
            logger.append(metric.evaluate(cfg["metric_name"]["test"]["Global"], inp, out), "test", bs)
