"""Multi-GPU host logic: filters shard by index, no data-path collective; the only exchange is
the final sum of the per-rank statistics vectors.  One process per GPU,
`torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is only the plumbing."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_filters: int, rank: int, world: int):
    """Contiguous index range [lo, hi) of rank `rank`: GPU g owns filters [g*N/G, (g+1)*N/G)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (n_filters * rank) // world
    hi = (n_filters * (rank + 1)) // world
    return lo, hi


def reduce_stats(stats, device=None, group=None):
    """Sums the per-rank [sum |e|^2, sum |e_xy|^2, n, n_bad] (kfpos_batch_error_stats) over all ranks
    with one all_reduce and returns (summed vector, rmse, rmse_xy).  Without an initialised process
    group it is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    v = torch.as_tensor(np.asarray(stats, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if device is not None:
            v = v.to(device)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        v = v.cpu()
    s = v.numpy()
    n = max(s[2], 1.0)
    return s, float(np.sqrt(s[0] / n)), float(np.sqrt(s[1] / n))
