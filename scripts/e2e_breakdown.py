"""Where the end-to-end step of bench.py spends its time: host seeding, PCIe copies, the pipelined call."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from rrtplanner_b200 import _lib, batch, worlds

W = H = 512; N = 5000; P = 4096
dev = torch.device("cuda", 0)
db = batch.DeviceBatch("star", W, H, N, 50.0, device=0).gen_worlds([worlds.world_seed(p) for p in range(P)])
og_pinned = torch.empty((P, W, H), dtype=torch.uint8, pin_memory=True); og_pinned.copy_(db.og); torch.cuda.synchronize()
t0 = time.perf_counter(); st = batch.seed_states(np.arange(P)); print("seed_states ms", 1e3 * (time.perf_counter() - t0))
d = torch.empty((P, W, H), dtype=torch.uint8, device=dev)
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(og_pinned, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D 1.07 GB ms", 1e3 * dt, "GB/s", og_pinned.numel() / dt / 1e9)
outp = torch.empty((P, N + 1, 2), dtype=torch.float64, pin_memory=True); src = torch.empty((P, N + 1, 2), dtype=torch.float64, device=dev)
for _ in range(2):
    t0 = time.perf_counter(); outp.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H 0.33 GB ms", 1e3 * dt, "GB/s", outp.numel() * 8 / dt / 1e9)
s = torch.cuda.Stream(); s2 = torch.cuda.Stream()
for _ in range(2):
    t0 = time.perf_counter()
    with torch.cuda.stream(s): d.copy_(og_pinned, non_blocking=True)
    with torch.cuda.stream(s2): outp.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("both directions together ms", 1e3 * dt)
ogs = og_pinned.numpy()
draws = np.random.default_rng(0).integers(0, 2, 1)
free0 = [np.argwhere(ogs[p] == 0)[[3, -3]] for p in range(8)]
starts = np.array([free0[p % 8][0] for p in range(P)]); goals = np.array([free0[p % 8][1] for p in range(P)])
# any valid start/goal will do for timing: take them from each world's own free cells
starts = np.array([np.argwhere(ogs[p, :64, :64] == 0)[0] for p in range(P)]); goals = np.array([np.argwhere(ogs[p, -64:, -64:] == 0)[-1] + (W - 64) for p in range(P)])
desc = batch.make_desc(np.arange(P), starts, goals)
out = (torch.empty((P, N + 1, 2), dtype=torch.int16, pin_memory=True).numpy(), torch.empty((P, N + 1), dtype=torch.float64, pin_memory=True).numpy(),
       torch.empty((P, N + 1), dtype=torch.int32, pin_memory=True).numpy(), torch.empty((P, _lib.STAT_COUNT), dtype=torch.int64, pin_memory=True).numpy(), None)
ctx = _lib.Context()
for chunk in [int(c) for c in sys.argv[1:]] or (0, 1036, 518):
    for _ in range(3):
        t0 = time.perf_counter(); ctx.plan_worlds(_lib.KIND_STAR, ogs, desc, N, 50.0, states=st, out=out, chunk=chunk); dt = time.perf_counter() - t0
    print("plan_worlds chunk", chunk, "ms", 1e3 * dt)
db.set_plans(desc); db.seed_samples(np.arange(P))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); db.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("device plan kernel ms", 1e3 * dt)
