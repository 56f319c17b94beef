"""Which rank trains which client.

Clients are dealt to ranks longest-processing-time-first on an estimated cost ``steps * (c0 + c1 * rate)``.  The
constants come from measured isolated train-step times on a B200 (profiles/step_census.md: ~0.53 ms at level e ...
~0.85 ms at level a -- a batch-10 step is latency bound, so the cost grows far slower than the rate^2 FLOP count).
Every rank computes the same assignment from the same inputs; nothing is communicated.
"""
from __future__ import annotations

from typing import List, Sequence


def client_cost(rate: float, n_steps: int, c0: float = 0.52, c1: float = 0.33) -> float:
    return n_steps * (c0 + c1 * rate)


def assign_clients(costs: Sequence[float], world: int) -> List[List[int]]:
    """LPT greedy: returns ``per_rank[r] = [client positions]`` (positions index the round's active-user list)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    for lst in out:
        lst.sort()
    return out
