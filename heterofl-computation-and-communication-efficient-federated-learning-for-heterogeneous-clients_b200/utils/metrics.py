"""Metrics: top-1 accuracy (%), perplexity = exp(CE), and the name table incl. ``Local-*`` / ``Global-*`` aliases
(reference ``src/metrics/metrics.py:7-44``).

The reference calls ``.item()`` on every metric every step (one host sync per step).  Here the training loop
accumulates ``(sum loss*n, sum correct, n)`` in a device tensor (:class:`DeviceMeter`) and reads it back once per
client / round; :class:`Metric` keeps the per-batch API for evaluation code and tests.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F


def accuracy(score: torch.Tensor, target: torch.Tensor, topk: int = 1) -> float:
    with torch.no_grad():
        pred = score.topk(topk, 1, True, True)[1]
        correct = pred.eq(target.view(-1, 1).expand_as(pred)).float().sum()
        return (correct * (100.0 / target.size(0))).item()


def perplexity(score: torch.Tensor, target: torch.Tensor) -> float:
    with torch.no_grad():
        return torch.exp(F.cross_entropy(score, target)).item()


class Metric:
    def __init__(self):
        loss = lambda inp, out: out["loss"].item()  # noqa: E731
        acc = lambda inp, out: accuracy(out["score"], inp["label"])  # noqa: E731
        ppl = lambda inp, out: (perplexity(out["score"], inp["label"]) if "score" in out  # noqa: E731
                                else math.exp(out["loss"].item()))
        self.metric = {}
        for pre in ("", "Local-", "Global-"):
            self.metric[pre + "Loss"] = loss
            self.metric[pre + "Accuracy"] = acc
            self.metric[pre + "Perplexity"] = ppl

    def evaluate(self, metric_names: List[str], inp: Dict, out: Dict) -> Dict[str, float]:
        return {n: self.metric[n](inp, out) for n in metric_names}


class DeviceMeter:
    """On-device accumulator ``[sum(loss*n), sum(correct), sum(n), sum(exp(loss)*n)]`` -- no host sync per step."""

    def __init__(self, device):
        self.buf = torch.zeros(4, dtype=torch.float64, device=device)

    def zero_(self):
        self.buf.zero_()

    def add(self, loss: torch.Tensor, n: int, score: torch.Tensor = None, target: torch.Tensor = None):
        with torch.no_grad():
            lf = loss.detach().double()
            self.buf[0] += lf * n
            self.buf[2] += n
            self.buf[3] += torch.exp(lf) * n
            if score is not None:
                self.buf[1] += (score.argmax(dim=1) == target).sum()

    def result(self, prefix: str = "", lm: bool = False) -> Dict[str, float]:
        s = self.buf.tolist()
        n = max(s[2], 1.0)
        out = {prefix + "Loss": s[0] / n}
        if lm:
            out[prefix + "Perplexity"] = s[3] / n
        else:
            out[prefix + "Accuracy"] = 100.0 * s[1] / n
        return out

    @property
    def count(self) -> float:
        return float(self.buf[2].item())
