"""Multi-GPU driver: one process per GPU, plans sharded by index, no collective in the loop.

Plans are independent (all state of the reference's ``plan()`` is local, rrt.py:487-494), so rank r
runs plans ``shard(nplans, r, world)`` on its own GPU with the single-GPU kernels.  The only
communication is the final gather of fixed-size per-plan records (trees, statistics) to one rank:
``torch.distributed.gather`` over NCCL (NVLink / NVSwitch) on GPUs, over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

from .batch import shard


def gather_records(local: Dict[str, np.ndarray], nplans: int, dst: int = 0, device=None, group=None) -> Optional[Dict[str, np.ndarray]]:
    """Gather per-plan arrays (leading dimension = this rank's shard, in shard order) to ``dst``.
    Returns the concatenated arrays (leading dimension nplans) on ``dst`` and None elsewhere."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = (nplans + world - 1) // world
    out = {} if rank == dst else None
    for name in sorted(local):
        a = np.ascontiguousarray(local[name])
        pad = np.zeros((per,) + a.shape[1:], dtype=a.dtype)
        pad[: a.shape[0]] = a
        t = torch.from_numpy(pad.view(np.uint8).reshape(per, -1))
        if device is not None:
            t = t.to(device)
        bucket = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bucket, dst=dst, group=group)
        if rank == dst:
            parts = []
            for r in range(world):
                m = len(shard(nplans, r, world))
                raw = bucket[r].cpu().numpy().reshape(-1)[: m * int(np.prod(a.shape[1:], dtype=np.int64)) * a.dtype.itemsize]
                parts.append(raw.view(a.dtype).reshape((m,) + a.shape[1:]))
            out[name] = np.concatenate(parts, axis=0)
    return out


def run_sharded(nplans: int, run_shard: Callable[[range], Dict[str, np.ndarray]], dst: int = 0, device=None, group=None):
    """Run ``run_shard(plan_indices)`` on every rank and gather the records on ``dst``."""
    import torch.distributed as dist

    ids = shard(nplans, dist.get_rank(group), dist.get_world_size(group))
    return gather_records(run_shard(ids), nplans, dst=dst, device=device, group=group)
