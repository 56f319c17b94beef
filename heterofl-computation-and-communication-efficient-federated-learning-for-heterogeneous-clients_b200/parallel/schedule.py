"""Which rank trains which client.

A round's cost is dominated by its widest clients (a level-a step moves/computes ~250x a level-e step), so clients
are dealt to ranks longest-processing-time-first on an estimated cost ``steps * (c0 + c1 * rate^2)`` -- ``c0`` models
the launch-latency floor that dominates narrow levels, ``c1 * rate^2`` the FLOP/byte growth of wide ones.
Every rank computes the same assignment from the same inputs; nothing is communicated.
"""
from __future__ import annotations

from typing import List, Sequence


def client_cost(rate: float, n_steps: int, c0: float = 1.0, c1: float = 3.0) -> float:
    return n_steps * (c0 + c1 * rate * rate)


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
