"""Multi-GPU sharding helpers (host logic only; one process per GPU, no collective on the data path).

Pairs / reads are independent, so the N>1 path is: partition by cost, run the single-GPU C-ABI on
each rank's shard, gather the per-rank results back into input order. Timing across ranks is the
max of the device times (bench.py does that with one all-reduce MAX outside the data path).
"""
import numpy as np


def partition(costs, world):
    """Longest-processing-time-first split of items with the given costs (e.g. Q*T cells) over
    `world` ranks. Returns a list of index arrays (ascending inside each rank); deterministic."""
    import heapq
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    heap = [(0.0, r) for r in range(world)]      # (load, rank): ties go to the lowest rank
    owner = np.empty(len(costs), dtype=np.int64)
    for i in order:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + float(costs[i]), r))
    return [np.nonzero(owner == r)[0] for r in range(world)]


def gather_in_order(local_idx, local_results, n_total, group=None):
    """All ranks get the list of results in input order. `local_results[k]` belongs to input item
    `local_idx[k]`. Uses torch.distributed object gather (results are small host objects)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, (list(map(int, local_idx)), list(local_results)), group=group)
    out = [None] * n_total
    for idx, res in bucket:
        for i, r in zip(idx, res):
            out[i] = r
    return out


def max_over_ranks(value, device=None, group=None):
    """Max of a scalar over ranks (the timing rule for multi-GPU numbers)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
