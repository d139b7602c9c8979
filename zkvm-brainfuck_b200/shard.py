"""Multi-GPU sharding of the proving hot path: one process per GPU, no data-path collective.

Trace matrices (chips) and whole proofs are independent units: every rank commits / proves the units the
plan assigns to it and only 8-word roots (and timing scalars) cross ranks, through `torch.distributed`
(NCCL on the GPU box, gloo in the CPU tests).  Sharding ONE commitment by rows (column shards -> row shards
all-to-all, then an all-gather of subtree caps) is designed in DESIGN.md §7 but not built in round 1.
"""
import numpy as np


def plan_units(costs, world_size):
    """Deterministic longest-processing-time assignment of work units to ranks.
    costs: list of per-unit costs (e.g. cells of a trace matrix).  Returns [rank of unit i]."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


def gather_roots(local_roots, dist=None):
    """All-gather the 8-word roots of every rank's units -> list (rank-major) of (n_r, 8) uint32 arrays."""
    local = np.ascontiguousarray(np.asarray(local_roots, np.uint32).reshape(-1, 8))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    import torch
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64))
    m = int(max(c.item() for c in counts))
    buf = torch.zeros((m, 8), dtype=torch.int64)
    buf[:local.shape[0]] = torch.from_numpy(local.astype(np.int64))
    out = [torch.zeros((m, 8), dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    return [o[:int(c.item())].numpy().astype(np.uint32) for o, c in zip(out, counts)]


def max_over_ranks(value, dist=None, device=None):
    """Device-time reduction used by bench.py: the job time is the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
