"""BASELINE.json's configurations at their FULL sizes on the GPU (cfg2: 1 Mi segments on a 2048 x 2048 grid; cfg3: 4096
plans of n = 5000 on 512 x 512 worlds), checked against the C oracle where it finishes in seconds and through
size-independent properties everywhere else.  (cfg5 at full size: tests/test_gpu_rewire.py::test_full_size_batch_properties.)"""
import numpy as np
import pytest

from oracle import c_oracle, rrt_oracle as O
from rrtplanner_b200 import _lib, batch, worlds

pytestmark = pytest.mark.gpu


def test_cfg2_one_million_segments_against_the_oracle():
    import torch
    S, nseg = 2048, 1 << 20
    db = batch.DeviceBatch("standard", S, S, 8).gen_worlds([worlds.world_seed(0)])
    og = db.og[0].cpu().numpy()
    segs = np.random.default_rng(0).integers(0, S, size=(nseg, 4)).astype(np.int32)     # bench.py's cfg2 segments
    d_segs = torch.from_numpy(segs).cuda()
    free = torch.empty(nseg, dtype=torch.uint8, device="cuda")
    cells = torch.empty(nseg, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(db.L.rrtk_collision_segments(db.bits.data_ptr(), S, S, d_segs.data_ptr(), None, nseg, free.data_ptr(), cells.data_ptr(), st))
    want_free, want_cells = c_oracle.collision_batch(og, segs)              # every one of the 2^20 segments (C: ~0.4 s)
    assert np.array_equal(free.cpu().numpy().astype(bool), want_free)
    assert np.array_equal(cells.cpu().numpy(), want_cells)
    # the clearance-field kernel gives the same two outputs
    clear = torch.empty((S, S), dtype=torch.uint8, device="cuda")
    scratch = torch.empty((2 * db.words,), dtype=torch.int32, device="cuda")
    _lib.check(db.L.rrtk_clearance_field(db.bits.data_ptr(), 1, S, S, 128, clear.data_ptr(), scratch.data_ptr(), st))
    free2, cells2 = torch.empty_like(free), torch.empty_like(cells)
    _lib.check(db.L.rrtk_collision_segments_cf(clear.data_ptr(), S, S, d_segs.data_ptr(), None, nseg, free2.data_ptr(), cells2.data_ptr(), st))
    assert torch.equal(free, free2) and torch.equal(cells, cells2)
    # and so does the walk on the eight directional fields (32 MB)
    clear8 = torch.empty((8, S, S), dtype=torch.uint8, device="cuda")
    _lib.check(db.L.rrtk_clearance_field_dir(db.bits.data_ptr(), 1, S, S, 255, clear8.data_ptr(), st))
    free2.zero_(); cells2.zero_()
    _lib.check(db.L.rrtk_collision_segments_cfd(clear8.data_ptr(), S, S, d_segs.data_ptr(), None, nseg, free2.data_ptr(), cells2.data_ptr(), st))
    assert torch.equal(free, free2) and torch.equal(cells, cells2)
    # and on the sixteen half-octant fields (64 MB)
    clear16 = torch.empty((16, S, S), dtype=torch.uint8, device="cuda")
    _lib.check(db.L.rrtk_clearance_field_dir16(db.bits.data_ptr(), 1, S, S, 255, clear16.data_ptr(), st))
    free2.zero_(); cells2.zero_()
    _lib.check(db.L.rrtk_collision_segments_cfd16(clear16.data_ptr(), S, S, d_segs.data_ptr(), None, nseg, free2.data_ptr(), cells2.data_ptr(), st))
    assert torch.equal(free, free2) and torch.equal(cells, cells2)
    del clear16
    # reversing a free segment keeps it free only if the reversed walk is free too: verdicts of both directions vs the oracle
    rev = np.ascontiguousarray(segs[: 1 << 17][:, [2, 3, 0, 1]])
    d_rev = torch.from_numpy(rev).cuda()
    _lib.check(db.L.rrtk_collision_segments(db.bits.data_ptr(), S, S, d_rev.data_ptr(), None, rev.shape[0], free2.data_ptr(), None, st))
    assert np.array_equal(free2[: rev.shape[0]].cpu().numpy().astype(bool), c_oracle.collision_batch(og, rev)[0])


def test_cfg3_4096_plans_properties_and_oracle_subset():
    import torch
    W = H = 512
    n, P = 5000, 4096
    db = batch.DeviceBatch("star", W, H, n, 50.0).gen_worlds([worlds.world_seed(p) for p in range(P)])
    pair = batch.DeviceBatch("star", W, H, 8)
    pair.bits, pair.rowcum = db.bits, db.rowcum
    pair.set_plans(batch.make_desc(np.arange(P), np.zeros((P, 2)), np.zeros((P, 2))))
    pair.seed_samples(2000 + np.arange(P))
    d = pair.samples.cpu().numpy().astype(np.int64)
    starts = d[:, 0]
    differs = (d[:, 1:] != starts[:, None]).any(axis=2)
    goals = d[np.arange(P), 1 + differs.argmax(axis=1)]
    db.set_plans(batch.make_desc(np.arange(P), starts, goals))
    db.seed_samples(np.arange(P))
    res = db.run().download()
    j, found, vgoal = res.stat("j"), res.stat("found"), res.stat("vgoal")
    top = j + found
    assert (j >= 1).all() and (j <= n).all() and ((vgoal == j) | (found == 0)).all() and found.mean() > 0.9
    rows = np.arange(n + 1)[None, :]
    live = (rows >= 1) & (rows < top[:, None])
    par = np.where(live, res.parent, 0)
    assert (res.parent[live] >= 0).all() and (par < np.maximum(rows, 1))[live | ~live].all()       # parents precede children (no rewiring)
    pts = res.pts.astype(np.int64)
    ppts = np.take_along_axis(pts, par[:, :, None].repeat(2, axis=2), axis=1)
    seg = np.sqrt(((pts - ppts) ** 2).sum(2).astype(np.float64))
    pc = np.take_along_axis(res.cost, par, axis=1)
    assert np.array_equal((pc + seg)[live].view(np.int64), res.cost[live].view(np.int64))           # cost = parent's cost + length, exactly
    assert (res.pts[~(rows < top[:, None])] == -32768).all() and np.isinf(res.cost[~(rows < top[:, None])]).all()
    # the sample stream accounts for every vertex: vertex v >= 1 of plan p is one of plan p's samples, in stream order
    smp = db.samples.cpu().numpy()
    for p in range(0, P, 256):
        jp = int(j[p])
        key = (smp[p].astype(np.int64) @ np.array([1, 1 << 20])).tolist()
        vk = (pts[p, 1:jp] @ np.array([1, 1 << 20])).tolist()
        i = 0
        for k in vk:                                   # the vertices are a subsequence of the stream (a cell can be drawn twice)
            while i < n and key[i] != k:
                i += 1
            assert i < n
            i += 1
        assert len(set(vk)) == len(vk)                 # `sampled` set: no cell twice among the vertices >= 1
    # every tree edge of every plan re-tested parent -> child by the stand-alone collision kernel (19.6 M segments)
    pi, vi = np.nonzero(live)
    segs = np.concatenate([ppts[pi, vi], pts[pi, vi]], axis=1).astype(np.int32)
    d_segs, d_w = torch.from_numpy(segs).cuda(), torch.from_numpy(pi.astype(np.int32)).cuda()
    ok = torch.empty(segs.shape[0], dtype=torch.uint8, device="cuda")
    _lib.check(db.L.rrtk_collision_segments(db.bits.data_ptr(), W, H, d_segs.data_ptr(), d_w.data_ptr(), segs.shape[0], ok.data_ptr(), None,
                                            torch.cuda.current_stream().cuda_stream))
    assert bool(ok.all().item())
    # bit-exact trees against the C oracle for a spread of plans
    ogs = db.og.cpu().numpy()
    for p in range(0, P, 256):
        wp, wc, wpar, st, _ = c_oracle.plan_raw("star", ogs[p], n, starts[p], goals[p], smp[p], 50.0)
        t = st["j"] + (1 if st["found"] else 0)
        assert int(j[p]) == st["j"] and int(vgoal[p]) == st["vgoal"]
        assert np.array_equal(res.pts[p, :t], wp[:t]) and np.array_equal(res.parent[p, :t], wpar[:t])
        assert np.array_equal(res.cost[p, :t].view(np.int64), wc[:t].view(np.int64))


def test_cfg4_1024_informed_pairs_on_one_world():
    """BASELINE cfg4 at full size: one 1024 x 1024 world, 1024 start/goal pairs, n = 20000, r = 50, r_goal = 5."""
    W = H = 1024
    n, P = 20000, 1024
    db = batch.DeviceBatch("informed", W, H, n, 50.0, 5.0).gen_worlds([worlds.world_seed(0)])
    og = db.og[0].cpu().numpy()
    pair = batch.DeviceBatch("star", W, H, 8)
    pair.bits, pair.rowcum = db.bits, db.rowcum
    pair.set_plans(batch.make_desc(np.zeros(P, int), np.zeros((P, 2)), np.zeros((P, 2))))
    pair.seed_samples(2000 + np.arange(P))
    d = pair.samples.cpu().numpy().astype(np.int64)
    starts = d[:, 0]
    differs = (d[:, 1:] != starts[:, None]).any(axis=2)
    goals = d[np.arange(P), 1 + differs.argmax(axis=1)]
    rots = np.stack([O.ellipse_rotation(a, b) for a, b in zip(starts, goals)])
    db.set_plans(batch.make_desc(np.zeros(P, int), starts, goals, rots))
    db.seed_samples(np.arange(P))
    u = np.random.default_rng(7).uniform(0, 1, size=(P, n, 2))
    balls = np.stack([np.sqrt(u[..., 0]) * np.cos(2 * np.pi * u[..., 1]), np.sqrt(u[..., 0]) * np.sin(2 * np.pi * u[..., 1])], axis=-1)
    db.set_balls_host(balls)
    res = db.run().download()
    j, found = res.stat("j"), res.stat("found")
    first, ell = res.stat("first_solution_iter"), res.stat("ellipse_iters")
    top = j + found
    assert found.all() and (first >= 0).mean() > 0.5                              # one open world: every goal is connected
    assert np.array_equal(ell, np.where(first >= 0, n - 1 - first, 0))             # every iteration after the first solution samples the ellipse
    rows = np.arange(n + 1)[None, :]
    live = (rows >= 1) & (rows < top[:, None])
    par = np.where(live, res.parent, 0)
    assert (res.parent[live] >= 0).all() and (par < np.maximum(rows, 1)).all()
    pts = res.pts.astype(np.int64)
    ppts = np.take_along_axis(pts, par[:, :, None].repeat(2, axis=2), axis=1)
    seg = np.sqrt(((pts - ppts) ** 2).sum(2).astype(np.float64))
    pc = np.take_along_axis(res.cost, par, axis=1)
    assert np.array_equal((pc + seg)[live].view(np.int64), res.cost[live].view(np.int64))
    # ellipse samples lie inside the grid; ellipse costs are recorded exactly for the plans that found a solution vertex, and
    # are at least the straight-line distance start -> goal (rrt.py:698-699: cost to a vertex near the goal + the rest)
    assert (pts[live] >= 0).all() and (pts[live] < W).all()
    for p in range(0, P, 64):
        c = res.ell_c[p][~np.isnan(res.ell_c[p])]
        assert (c.size > 0) == (first[p] >= 0)
        assert (c >= np.hypot(*(starts[p] - goals[p])) - 1e-9).all()
    smp = db.samples.cpu().numpy()
    for p in (0, 517):                                                             # bit-exact against the C oracle (about 1 s per plan)
        wp, wc, wpar, st, well = c_oracle.plan_raw("informed", og, n, starts[p], goals[p], smp[p], 50.0, 5.0, balls[p], rots[p])
        t = st["j"] + (1 if st["found"] else 0)
        assert int(j[p]) == st["j"] and int(first[p]) == st["first_solution_iter"]
        assert np.array_equal(res.pts[p, :t], wp[:t]) and np.array_equal(res.parent[p, :t], wpar[:t])
        assert np.array_equal(res.cost[p, :t].view(np.int64), wc[:t].view(np.int64))
