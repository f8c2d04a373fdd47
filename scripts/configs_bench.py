"""Timings of the BASELINE configurations that are not the bench line (cfg1 latency through the class API,
cfg3 with RRTStandard, cfg4 = RRTStarInformed on one 1024^2 world).  Prints one JSON object per line."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from rrtplanner_b200 import batch, worlds, rrt


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def pairs_on_device(db, P, seed0):
    pdb = batch.DeviceBatch("star", db.W, db.H, 8, device=0)
    pdb.bits, pdb.rowcum = db.bits, db.rowcum
    pdb.set_plans(db_desc(db, P, np.zeros((P, 2)), np.zeros((P, 2))))
    pdb.seed_samples(seed0 + np.arange(P))
    d = pdb.samples.cpu().numpy().astype(np.int64)
    s = d[:, 0]
    differs = (d[:, 1:] != s[:, None]).any(axis=2)
    return s, d[np.arange(P), 1 + differs.argmax(axis=1)]


def db_desc(db, P, starts, goals, rots=None):
    wid = np.arange(P) % db.bits.shape[0]
    return batch.make_desc(wid, starts, goals, rots)


# cfg1: 256^2, n=1000, r=50, one plan through the drop-in class (includes graph construction on the host)
og = worlds.perlin_occupancygrid(256, 256, seed=worlds.world_seed(0))
xs, xg = worlds.start_goal(og, 0)
pl = rrt.RRTStar(og, 1000, 50.0, pbar=False, seed=0)
pl.plan(xs, xg)
t0 = time.perf_counter()
for _ in range(5):
    T, gv = pl.plan(xs, xg)
dt = (time.perf_counter() - t0) / 5
print(json.dumps({"config": "cfg1 RRTStar 256x256 n=1000 r=50, class API incl. networkx graph", "s_per_plan": dt, "plans_per_s": 1 / dt, "nodes": T.number_of_nodes()}))

# cfg3 shape with RRTStandard
P = 2960
db = batch.DeviceBatch("standard", 512, 512, 5000, device=0).gen_worlds([worlds.world_seed(p) for p in range(P)])
s, g = pairs_on_device(db, P, 2000)
db.set_plans(db_desc(db, P, s, g)); db.seed_samples(np.arange(P))
ms = timed(lambda: db.run())
print(json.dumps({"config": "cfg3 shape, RRTStandard, %d plans" % P, "ms": ms, "plans_per_s": P / ms * 1e3}))

# cfg4: one 1024^2 world, 1024 pairs, n=20000, r=50, r_goal=5
P, n = 1024, 20000
db = batch.DeviceBatch("informed", 1024, 1024, n, 50.0, 5.0, device=0).gen_worlds([worlds.world_seed(0)])
s, g = pairs_on_device(db, P, 2000)
from rrtplanner_b200.rrt import RRTStarInformed
og1 = db.og[0].cpu().numpy()
helper = RRTStarInformed(og1, 8, 50.0, 5.0, pbar=False)
rots = np.stack([np.asarray(helper.rotation_to_world_frame(a, b), dtype=np.float64) for a, b in zip(s, g)])
db.set_plans(db_desc(db, P, s, g, rots)); db.seed_samples(np.arange(P))
rng = np.random.default_rng(7)
u = rng.uniform(0, 1, size=(P, n, 2))
balls = np.stack([np.sqrt(u[..., 0]) * np.cos(2 * np.pi * u[..., 1]), np.sqrt(u[..., 0]) * np.sin(2 * np.pi * u[..., 1])], axis=-1)
db.set_balls_host(balls)
ms = timed(lambda: db.run(), reps=2)
st = db.out["stats"].cpu().numpy()
print(json.dumps({"config": "cfg4 RRTStarInformed, one 1024x1024 world, %d pairs, n=%d" % (P, n), "ms": ms, "plans_per_s": P / ms * 1e3,
                  "found_frac": float(st[:, 2].mean()), "mean_first_solution_iter": float(st[st[:, 5] >= 0, 5].mean()),
                  "ellipse_iter_frac": float(st[:, 6].mean() / n), "blocks_per_sm": db.footprint()[1]}))
