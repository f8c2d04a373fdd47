"""Experiment: per-phase cycle counts of the plan kernel (library built with -DRRTK_PHASE_CLOCKS, RRTK_LIB=...)."""
import sys, numpy as np
sys.path.insert(0, '.')
from rrtplanner_b200 import _lib, batch, worlds
import torch
P = int(sys.argv[1]) if len(sys.argv) > 1 else 740
T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W = H = 512; n = 5000
db = batch.DeviceBatch("star", W, H, n, 50.0, device=0, threads=T).gen_worlds([worlds.world_seed(p) for p in range(P)])
ogs = db.og[:1].cpu().numpy()
pair_db = batch.DeviceBatch("star", W, H, 8, device=0)
pair_db.bits, pair_db.rowcum = db.bits, db.rowcum
pair_db.set_plans(batch.make_desc(np.arange(P), np.zeros((P, 2)), np.zeros((P, 2))))
pair_db.seed_samples(2000 + np.arange(P))
d = pair_db.samples.cpu().numpy().astype(np.int64)
starts = d[:, 0]; differs = (d[:, 1:] != starts[:, None]).any(axis=2); goals = d[np.arange(P), 1 + differs.argmax(axis=1)]
db.set_plans(batch.make_desc(np.arange(P), starts, goals)); db.seed_samples(np.arange(P))
for _ in range(3): db.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); db.run(); e1.record(); torch.cuda.synchronize()
st = db.out["stats"].cpu().numpy()
S = {nm: st[:, i].astype(np.float64) for i, nm in enumerate(_lib.STAT_NAMES)}
names = list(_lib.STAT_NAMES)
scan, owner, commit, ownwork, rounds = (st[:, names.index(k)].astype(np.float64) for k in ("reserved0", "reserved1", "ellipse_iters", "first_solution_iter", "ring_members"))
tot = scan + owner + commit
print(f"P={P} T={T or 'default'} kernel {e0.elapsed_time(e1):.2f} ms; rounds/plan {rounds.mean():.0f}; cycles/round: scan {np.mean(scan/rounds):.0f} owner {np.mean(owner/rounds):.0f} (warp 0 busy {np.mean(ownwork/rounds):.0f}) commit {np.mean(commit/rounds):.0f}; total/plan {tot.mean()/1e6:.2f} Mcycles")
print(f"   per-plan cycles: mean {tot.mean()/1e6:.2f}M  p50 {np.percentile(tot,50)/1e6:.2f}M  p90 {np.percentile(tot,90)/1e6:.2f}M  max {tot.max()/1e6:.2f}M; rounds p50 {np.percentile(rounds,50):.0f} max {rounds.max():.0f}; kernel = {e0.elapsed_time(e1)*1.965e3/1e3:.2f} Mcycles at 1965 MHz")
