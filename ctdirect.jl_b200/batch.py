"""Batched MPC scenarios sharded over the GPUs of one node (BASELINE config 5).

Independent instances of the same DOCP (different parameters / initial states) are split into
contiguous blocks, one block per rank (one process per GPU).  The data path needs no collective; the only
exchange is one all-gather of the per-instance (objective, constraint violation) pairs, 2*B doubles per
evaluation, over NCCL (GPU tensors) or gloo (CPU tensors, used by the CPU tests).
The reference has no batching at all (every scenario would be its own DOCP)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_instances: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of instances owned by `rank`; blocks differ by at most one instance."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(n_instances, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_norms(obj, viol, n_instances: int, group=None):
    """All-gather the local (objective, violation) vectors into two global vectors of n_instances entries,
    in instance order, on every rank.  `obj` and `viol` are torch tensors (CUDA -> NCCL, CPU -> gloo)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return obj.clone(), viol.clone()
    world = dist.get_world_size(group)
    sizes = [shard_range(n_instances, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    local = torch.zeros((2, nmax), dtype=torch.float64, device=obj.device)
    local[0, :obj.numel()] = obj
    local[1, :viol.numel()] = viol
    out = torch.empty((world * 2, nmax), dtype=torch.float64, device=obj.device)
    dist.all_gather_into_tensor(out, local, group=group)   # concatenation along dim 0 (NCCL and gloo)
    out = out.view(world, 2, nmax)
    g_obj = torch.cat([out[r, 0, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    g_viol = torch.cat([out[r, 1, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    return g_obj, g_viol


class ShardedScenarios:
    """The local shard of a batch of scenarios: a batched DOCP handle for instances [lo, hi)."""

    def __init__(self, ocp, n_instances: int, params: np.ndarray, grid_size: int = 1000, scheme: str = "midpoint",
                 rank: int = 0, world_size: int = 1, device: Optional[int] = None, **kwargs):
        from .docp import DOCP
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim != 2 or params.shape[0] != n_instances:
            raise ValueError("params must have one row per instance")
        self.n_instances, self.rank, self.world_size = n_instances, rank, world_size
        self.lo, self.hi = shard_range(n_instances, world_size, rank)
        self.docp = DOCP(ocp, grid_size=grid_size, scheme=scheme, params=params[self.lo:self.hi],
                         batch=self.hi - self.lo, device=rank if device is None else device, **kwargs)

    def norms(self, x, group=None):
        """Local evaluation of (objective, violation) followed by the all-gather: global vectors."""
        obj, viol = self.docp.norms(x)
        return gather_norms(obj, viol, self.n_instances, group)
