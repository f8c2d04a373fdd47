"""Host-side multi-GPU logic on CPU: plan sharding and the final gather, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rrtplanner_b200 import batch, multigpu


def test_shard_partitions_exactly():
    for nplans in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            parts = [list(batch.shard(nplans, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(nplans))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= max(1, (nplans + world - 1) // world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nplans, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fake_run(ids):   # stands in for the GPU kernels: deterministic per-plan records
        ids = np.asarray(list(ids), dtype=np.int64)
        return {"stats": np.stack([ids, ids * 3, ids % 5], axis=1).reshape(-1, 3),
                "cost": (ids[:, None] * 0.5 + np.arange(4)[None]).astype(np.float64),
                "pts": np.stack([ids.astype(np.int16), (ids * 2).astype(np.int16)], axis=1)[:, None, :].repeat(3, axis=1)}

    got = multigpu.run_sharded(nplans, fake_run, dst=0)
    # the device-resident form (bench.py's strong-scaling leg gathers path records with it over NCCL): here CPU tensors / gloo
    mine = {k: torch.from_numpy(v) for k, v in fake_run(batch.shard(nplans, rank, world)).items()}
    got_t = multigpu.gather_tensors(mine, nplans, dst=0)
    if rank == 0:
        want = fake_run(range(nplans))
        ok = all(np.array_equal(got[k], want[k]) and got[k].dtype == want[k].dtype for k in want)
        ok = ok and all(np.array_equal(got_t[k].numpy(), want[k]) and got_t[k].numpy().dtype == want[k].dtype for k in want)
        q.put(ok)
    else:
        assert got is None and got_t is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nplans", [9, 64, 1])
def test_gather_world_size_2_gloo(nplans):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nplans, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
