"""Bytes moved by one round's aggregation / distribution and the achieved rate against the link / memory roofline.

BASELINE.json asks for "aggregate + distribute GB/s vs the NVLink roofline".  The fused aggregation kernel
(``csrc/fed_kernels.cu``, ``parallel/peer.py``) reads every active client's fp32 result arena once (over NVLink when the
client trained on a peer) and writes each coordinate of the new ``W_g`` into every replica; distribution reads ``W_g`` and
writes fp32 master + bf16 compute copy (+ zeroed momentum) of each starting client.  These phases are latency-bound
(~0.1-1.5 ms for 0.1-1.4 GB), so the fraction of the roofline is reported for honesty, not as a target.
"""
from __future__ import annotations

from typing import Dict, Sequence

NVLINK5_GBS_PER_DIR = 900.0        # per GPU, per direction (nominal)


def aggregate_traffic(level_sizes: Sequence[int], client_levels: Sequence[int], global_size: int, world: int,
                      agg_ms: float, hbm_gbs: float = 0.0) -> Dict[str, float]:
    """``level_sizes[l]`` = elements of a level-l client arena; ``client_levels`` = level of every aggregated client."""
    pulled = 4.0 * sum(int(level_sizes[l]) for l in client_levels)            # fp32 client arenas read once
    published = 4.0 * int(global_size) * max(int(world), 1)                    # every replica of W_g written once
    total = pulled + published
    out = {"client_bytes": pulled, "publish_bytes": published, "ms": float(agg_ms)}
    if agg_ms > 0:
        gbs = total / (agg_ms * 1e-3) / 1e9
        out["gbs"] = gbs
        if world > 1:
            # remote share of the traffic: (world-1)/world of the pulls and of the publishes cross NVLink
            remote = total * (world - 1) / world
            out["nvlink_gbs"] = remote / (agg_ms * 1e-3) / 1e9
            out["nvlink_roofline_frac"] = out["nvlink_gbs"] / (NVLINK5_GBS_PER_DIR * world)
        elif hbm_gbs:
            out["hbm_roofline_frac"] = gbs / hbm_gbs
    return out
