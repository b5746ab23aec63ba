"""Tile -> rank scheduling helpers (second half): which rank processes which tiles when tiles have
different costs. Host-side arithmetic only; same results as the reference's ``split_across_nodes``
(utils/distributed.py:240-309) and bucket assignment (utils/bucket_assignment.py:4-43), which it restates with a heap
(O(n log k) instead of an argmin per element) -- ties go to the lowest bucket index, as ``np.argmin`` does.

The multi-GPU merge itself uses contiguous row-major ranges (distributed/ragged.py: neighbours must overlap spatially for
the halo to be a single hop); these helpers are for the stage in front of it -- deciding which rank runs the MODEL on which
tiles when the per-tile cost is not uniform (e.g. tiles that are mostly border)."""
from __future__ import annotations

import heapq
from typing import List, Optional, Sequence

import numpy as np

__all__ = ["naive_bucket_assignment", "filler_bucket_assignment", "compute_bucket_imbalance_score", "split_across_nodes"]


def naive_bucket_assignment(costs: np.ndarray, num_buckets: int) -> np.ndarray:
    """Round-robin by POSITION IN THE SORTED ORDER'S INDEX, exactly as the reference computes it (argsort(costs) % k)."""
    return np.argsort(np.asarray(costs)) % num_buckets


def filler_bucket_assignment(costs: np.ndarray, num_buckets: int) -> np.ndarray:
    """Longest-processing-time-first: elements in decreasing cost order, each to the currently least loaded bucket."""
    costs = np.asarray(costs)
    order = np.argsort(-costs)
    heap = [(0.0, b) for b in range(num_buckets)]  # already a valid heap; ties pop the lowest bucket index first
    assignment = np.zeros(len(costs), dtype=int)
    for idx in order:
        load, bucket = heapq.heappop(heap)
        assignment[idx] = bucket
        heapq.heappush(heap, (load + float(costs[idx]), bucket))
    return assignment


def compute_bucket_imbalance_score(costs, assignment) -> float:
    costs, assignment = np.asarray(costs), np.asarray(assignment)
    return float(np.std([np.sum(costs[assignment == b]) for b in np.unique(assignment)]))


def split_across_nodes(collection: Sequence, world_size: Optional[int] = None, local_rank: Optional[int] = None,
                       cost: Optional[Sequence] = None, method: str = "optimal") -> List:
    """The part of ``collection`` rank ``local_rank`` of ``world_size`` processes: contiguous ``linspace`` slices without
    ``cost``; with it, the items whose bucket (``method`` "optimal" = filler, "naive") is this rank, in collection order."""
    if world_size is None or local_rank is None:
        import torch.distributed as dist

        ready = dist.is_available() and dist.is_initialized()
        world_size = (dist.get_world_size() if ready else 1) if world_size is None else world_size
        local_rank = (dist.get_rank() if ready else 0) if local_rank is None else local_rank
    if world_size <= 1:
        return collection
    if cost is None:
        edges = np.linspace(0, len(collection), int(world_size + 1), dtype=int)
        return collection[edges[local_rank] : edges[local_rank + 1]]
    if len(cost) != len(collection):
        raise RuntimeError("cost and collection must have the same length")
    method_fn = {"optimal": filler_bucket_assignment, "naive": naive_bucket_assignment}[method]
    mine = method_fn(np.asarray(cost), world_size) == local_rank
    if isinstance(collection, np.ndarray):
        return collection[mine]
    return [item for item, pick in zip(collection, mine) if pick]
