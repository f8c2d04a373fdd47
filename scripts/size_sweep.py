"""Experiment: plans/s of the bucket and the scan form of K7 over tree sizes n (512 x 512 worlds, r = 50, RRT*): where should the
dispatch switch?  RRTK_PLAN_IMPL=grid forces the bucket form for any size that fits, =scan forbids it."""
import os, subprocess, sys
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ".")
    import torch
    from rrtplanner_b200 import batch, worlds
    n, P = int(sys.argv[2]), int(sys.argv[3])
    W = H = 512
    db = batch.DeviceBatch("star", W, H, n, 50.0).gen_worlds([worlds.world_seed(p) for p in range(P)])
    og = db.og
    pair = batch.DeviceBatch("star", W, H, 8)
    pair.bits, pair.rowcum = db.bits, db.rowcum
    pair.set_plans(batch.make_desc(np.arange(P), np.zeros((P, 2)), np.zeros((P, 2))))
    pair.seed_samples(2000 + np.arange(P))
    d = pair.samples.cpu().numpy().astype(np.int64)
    starts = d[:, 0]; differs = (d[:, 1:] != starts[:, None]).any(axis=2); goals = d[np.arange(P), 1 + differs.argmax(axis=1)]
    db.set_plans(batch.make_desc(np.arange(P), starts, goals)); db.seed_samples(np.arange(P))
    for _ in range(2): db.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); db.run(); db.run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print("n %5d plans %5d impl %-5s kernel %-4s %8.2f ms %9.0f plans/s  blocks/SM %d" % (
        n, P, os.environ.get("RRTK_PLAN_IMPL", "auto"), db.L.rrtk_plan_kernel(db.kind, W, H, n, 0).decode(), ms, P / ms * 1e3, db.footprint()[1]))
else:
    for n in (1000, 1500, 2048, 3000, 5000, 5500, 6500, 8000):
        for impl in ("scan", "grid"):
            env = dict(os.environ, RRTK_PLAN_IMPL=impl)
            subprocess.run([sys.executable, __file__, "one", str(n), str(4096 if n <= 5000 else 2048)], env=env)
