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


def gather_tensors(local: Dict[str, "torch.Tensor"], nplans: int, dst: int = 0, group=None):
    """Device-resident form of gather_records: per-plan torch tensors (leading dimension = this rank's shard) are
    gathered to ``dst`` with one ``torch.distributed.gather`` per field on the tensors' own device (NCCL over NVLink for
    CUDA tensors, gloo for CPU tensors) -- no host staging.  Returns the concatenated tensors on ``dst``, None elsewhere.
    Shards are padded to the common size ceil(nplans / world); the padding is dropped on ``dst``."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = (nplans + world - 1) // world
    out = {} if rank == dst else None
    for name in sorted(local):
        a = local[name].contiguous()
        tail, dtype = tuple(a.shape[1:]), a.dtype
        m, rowbytes = a.shape[0], int(np.prod(tail, dtype=np.int64)) * a.element_size()
        if m == per:
            raw = a.reshape(m, -1).view(torch.uint8)                     # bytes: every backend moves uint8
        else:                                                            # short (or empty) shard: pad to the common size
            raw = torch.zeros((per, rowbytes), dtype=torch.uint8, device=a.device)
            if m:
                raw[:m] = a.reshape(m, -1).view(torch.uint8)
        if rank == dst:
            whole = torch.empty((world * per, raw.shape[1]), dtype=torch.uint8, device=raw.device)
            bucket = list(whole.split(per, dim=0))
        else:
            whole, bucket = None, None
        dist.gather(raw, bucket, dst=dst, group=group)
        if rank == dst:
            # shards are contiguous index ranges of `per` plans each (batch.shard), so only the tail is padding
            out[name] = whole[:nplans].view(dtype).reshape((nplans,) + tail)
    return out


def run_sharded(nplans: int, run_shard: Callable[[range], Dict[str, np.ndarray]], dst: int = 0, device=None, group=None):
    """Run ``run_shard(plan_indices)`` on every rank and gather the records on ``dst``."""
    import torch.distributed as dist

    ids = shard(nplans, dist.get_rank(group), dist.get_world_size(group))
    return gather_records(run_shard(ids), nplans, dst=dst, device=device, group=group)
