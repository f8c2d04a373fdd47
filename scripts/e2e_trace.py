"""Experiment: where the end-to-end call of bench.py (cfg3, packed grids in, path records out) spends its time.
RRTK_PIPE_TRACE=1 makes rrtk_ctx_plan_worlds2 print the device-side time stamps of every chunk."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import bench
from rrtplanner_b200 import _lib, batch

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ids = np.arange(P)
db, desc, _ = bench.cfg3_batch(0, ids)[:3]
bits_pinned = torch.empty((P, db.words), dtype=torch.int32, pin_memory=True); bits_pinned.copy_(db.bits); torch.cuda.synchronize()
bits_host = bits_pinned.numpy().view(np.uint32)
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
cap = 256
out = {"stats": pin((P, _lib.STAT_COUNT), torch.int64), "path": pin((P, cap), torch.int32), "xy": pin((P, cap, 2), torch.int16),
       "len": pin((P,), torch.int32), "path_cost": pin((P,), torch.float64)}
ctx = _lib.Context()
t0 = time.perf_counter(); st = batch.seed_states(ids); print("seed_states ms %.2f" % (1e3 * (time.perf_counter() - t0)))
for rep in range(4):
    if rep == 3: os.environ["RRTK_PIPE_TRACE"] = "1"
    t0 = time.perf_counter()
    ctx.plan_worlds2(_lib.KIND_STAR, bits_host, 512, 512, desc, 5000, 50.0, states=st, bits=True, trees=False, paths=True, path_cap=cap, out=out, chunk=chunk)
    print("call ms %.2f" % (1e3 * (time.perf_counter() - t0)), flush=True)
