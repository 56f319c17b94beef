"""Run logger: sample-weighted running means, per-round history, optional TensorBoard mirror.

Interface and *instance attribute names* (``log_path, writer, tracker, counter, mean, history, iterator``) match
reference ``src/logger.py:8-88`` because the logger object is pickled into every checkpoint under the class path
``logger.Logger`` (see :mod:`hfl_b200.ckpt.compat`) and ``history[tag/metric]`` -- one value per round -- is what
post-processing plots.
"""
from __future__ import annotations

from collections import defaultdict
from collections.abc import Iterable
from numbers import Number


def _make_writer(path):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(path)
    except Exception:
        return None


class Logger:
    def __init__(self, log_path, tensorboard: bool = True):
        self.log_path = log_path
        self.writer = None
        self.tracker = defaultdict(int)
        self.counter = defaultdict(int)
        self.mean = defaultdict(int)
        self.history = defaultdict(list)
        self.iterator = defaultdict(int)
        self._tb = tensorboard

    def safe(self, write: bool) -> None:
        """``safe(True)`` opens the event writer for a round; ``safe(False)`` closes it and appends the round's
        means to ``history``."""
        if write:
            if getattr(self, "_tb", True) and self.log_path:
                self.writer = _make_writer(self.log_path)
        else:
            if self.writer is not None:
                self.writer.close()
                self.writer = None
            for name in self.mean:
                self.history[name].append(self.mean[name])

    def reset(self) -> None:
        self.tracker = defaultdict(int)
        self.counter = defaultdict(int)
        self.mean = defaultdict(int)

    def append(self, result, tag, n=1, mean=True) -> None:
        for k, val in result.items():
            name = "{}/{}".format(tag, k)
            self.tracker[name] = val
            if not mean:
                continue
            if isinstance(val, Number):
                self.counter[name] += n
                c = self.counter[name]
                self.mean[name] = ((c - n) * self.mean[name] + n * val) / c
            elif isinstance(val, Iterable):
                if name not in self.mean:
                    self.counter[name] = [0 for _ in val]
                    self.mean[name] = [0 for _ in val]
                ns = list(n) if isinstance(n, Iterable) else [n] * len(val)
                for i, vi in enumerate(val):
                    self.counter[name][i] += ns[i]
                    c = self.counter[name][i]
                    self.mean[name][i] = ((c - ns[i]) * self.mean[name][i] + ns[i] * vi) / c
            else:
                raise ValueError("Not valid data type")

    def write(self, tag, metric_names, quiet: bool = False) -> str:
        parts = []
        for k in metric_names:
            name = "{}/{}".format(tag, k)
            m = self.mean[name]
            if isinstance(m, Number):
                parts.append("{}: {:.4f}".format(k, m))
                scalar = m
            elif isinstance(m, Iterable):
                parts.append("{}: {}".format(k, tuple(m)))
                scalar = tuple(m)[0]
            else:
                raise ValueError("Not valid data type")
            if self.writer is not None:
                self.iterator[name] += 1
                self.writer.add_scalar(name, scalar, self.iterator[name])
        info_name = "{}/info".format(tag)
        info = list(self.tracker[info_name]) if isinstance(self.tracker[info_name], list) else []
        info[2:2] = parts
        line = "  ".join(info)
        if not quiet:
            print(line)
        if self.writer is not None:
            self.iterator[info_name] += 1
            self.writer.add_text(info_name, line, self.iterator[info_name])
        return line

    def flush(self) -> None:
        if self.writer is not None:
            self.writer.flush()


# class path under which instances are pickled (checkpoint compatibility with the reference)
Logger.__module__ = "logger"
