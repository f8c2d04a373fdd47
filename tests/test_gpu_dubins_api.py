"""GPU tests of the host mirror of K8: ``RRTStar(rewire="rrtstar")``, ``RRTDubins`` / ``RRTStarDubins`` and the Dubins
primitive module, called the way a user of the reference's classes would call them and compared with the
specification oracle (oracle/rewire_oracle.c; these planners do not exist in the reference -- parity UNPINNED)."""
import networkx as nx
import numpy as np
import pytest

import rrtplanner_b200 as R
from oracle import rewire_oracle as O2
from rrtplanner_b200 import _lib, batch, worlds

pytestmark = pytest.mark.gpu


def _og(seed, W=128, H=112):
    return worlds.perlin_occupancygrid(W, H, seed=seed)


def test_rrtstar_rewire_mode_matches_the_specification_from_the_seed_alone():
    og = _og(3)
    free = np.argwhere(og == 0)
    xs, xg = free[5], free[-5]
    n, r, seed = 800, 25.0, 11
    p = R.RRTStar(og, n, r, pbar=False, seed=seed, rewire="rrtstar")
    T, gv = p.plan(xs, xg)
    smp = free[np.random.default_rng(seed).integers(0, len(free), n)]
    want = O2.plan("euclid", og, n, [*xs, 0], [*xg, 0], np.concatenate([smp, np.zeros((n, 1), dtype=np.int64)], axis=1),
                   star=True, rewire=True, r_rewire=r)
    st = want["stats"]
    assert p.last_stats["rewires"] == st["rewires"] > 0 and p.last_stats["j"] == st["j"]
    assert int(gv) == (st["vgoal"] if st["found"] else 0)
    top = st["j"] + (1 if st["found"] else 0)
    for v in range(1, top):
        par = int(want["parent"][v])
        assert T.has_edge(par, v)
        assert T.edges[par, v]["cost"] == want["cost"][v] and T.edges[par, v]["dist"] == want["elen"][v]
        assert np.array_equal(T.nodes[v]["pt"], want["pts"][v])
    assert T.number_of_edges() == top - 1
    # rewired trees have edges from younger to older vertices; the reference-mode tree never does
    assert any(a > b for a, b in T.edges if b != gv)
    T0, _ = R.RRTStar(og, n, r, pbar=False, seed=seed).plan(xs, xg)
    assert not any(a > b for a, b in T0.edges)
    path = p.route2gv(T, gv)
    assert path[0] == 0 and path[-1] == gv
    with pytest.raises(ValueError):
        R.RRTStar(og, 10, 5.0, rewire="yes")


@pytest.mark.parametrize("cls", ["rrt", "star", "star_norewire"])
def test_dubins_planners_match_the_specification_from_the_seed_alone(cls):
    og = _og(4)
    free = np.argwhere(og == 0)
    n, r, rho, nh, ds, seed = 700, 30.0, 4.0, 16, 1.0, 5
    xs, xg = np.array([*free[9], 3]), np.array([*free[-9], 12])
    if cls == "rrt":
        p = R.RRTDubins(og, n, rho, nheadings=nh, ds=ds, pbar=False, seed=seed)
    else:
        p = R.RRTStarDubins(og, n, r, rho, nheadings=nh, ds=ds, pbar=False, seed=seed, rewire=cls == "star")
    T, gv = p.plan(xs, xg)
    g = np.random.default_rng(seed)
    cells = free[g.integers(0, len(free), n)]
    heads = g.integers(0, nh, n)
    want = O2.plan("dubins", og, n, xs, xg, np.concatenate([cells, heads[:, None]], axis=1), star=cls != "rrt",
                   rewire=cls == "star", r_rewire=r if cls != "rrt" else 0.0, nh=nh, rho=rho, ds=ds)
    st = want["stats"]
    assert st["j"] > 50
    assert int(gv) == (st["vgoal"] if st["found"] else 0)
    assert T.number_of_nodes() == (n + 1 if st["found"] else n)        # same node layout as the reference's planners
    top = st["j"] + (1 if st["found"] else 0)
    for v in range(1, top):
        par = int(want["parent"][v])
        assert T.edges[par, v]["cost"] == want["cost"][v] and T.edges[par, v]["dist"] == want["elen"][v]
        assert T.nodes[v]["heading"] == want["head"][v] and np.array_equal(T.nodes[v]["pt"], want["pts"][v])
    if cls == "star":
        assert p.last_stats["rewires"] == st["rewires"] > 0
    if st["found"]:
        path = p.route2gv(T, gv)
        assert path[0] == 0 and path[-1] == gv
        poses = p.path_points(T, path)
        assert poses.shape[1] == 3 and np.allclose(poses[0, :2], xs[:2])
        cx, cy = np.floor(poses[:, 0] + 0.5).astype(int), np.floor(poses[:, 1] + 0.5).astype(int)
        assert (og[cx, cy] == 0).all()                                   # the flown path never touches an obstacle
        step = np.hypot(np.diff(poses[:, 0]), np.diff(poses[:, 1]))
        assert step.max() <= ds + 1e-9
        length = sum(T.edges[a, b]["dist"] for a, b in zip(path[:-1], path[1:]))
        assert abs(length - float(T.edges[path[-2], path[-1]]["cost"])) < 1e-9 * max(1.0, length)


def test_dubins_primitive_module():
    q0, q1 = np.array([10, 10, 0]), np.array([40, 25, 4])
    word, tpq, ln = R.dubins_path(q0, q1, 5.0, 16)
    w2, t2, l2 = O2.dubins(np.concatenate([q0, q1])[None], 16, 5.0)
    assert word == w2[0] and np.array_equal(tpq, t2[0]) and ln == l2[0]
    assert R.dubins.DUBINS_WORDS[word] in ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")
    assert R.dubins_length(q0, q1, 5.0, 16) == ln
    pts = R.dubins_points(q0, q1, 5.0, 16, 0.5)
    assert pts.shape == (int(np.floor(ln / 0.5)) + 1, 3)
    assert np.array_equal(pts, O2.dubins_points(np.concatenate([q0, q1]), 16, 5.0, np.arange(pts.shape[0]) * 0.5))
    og = np.zeros((64, 64))
    assert R.dubins_collisionfree(og, q0, q1, 5.0, 16, 1.0)
    og[20:30, 0:40] = 1
    assert not R.dubins_collisionfree(og, q0, q1, 5.0, 16, 1.0)
    # arrays of pairs
    rng = np.random.default_rng(0)
    a = np.stack([rng.integers(0, 64, 50), rng.integers(0, 64, 50), rng.integers(0, 16, 50)], axis=1)
    b = np.stack([rng.integers(0, 64, 50), rng.integers(0, 64, 50), rng.integers(0, 16, 50)], axis=1)
    _, _, lens = R.dubins_path(a, b, 5.0, 16)
    assert np.array_equal(lens, O2.dubins(np.concatenate([a, b], axis=1), 16, 5.0)[2])
    assert np.array_equal(R.dubins_collisionfree(og, a, b, 5.0, 16, 1.0), O2.dubins_free(og, np.concatenate([a, b], axis=1), 16, 5.0, 1.0))


def test_bad_configurations_raise():
    og = np.zeros((32, 32))
    p = R.RRTStarDubins(og, 50, 10.0, 3.0, pbar=False)
    with pytest.raises(ValueError):
        p.plan(np.array([1, 1]), np.array([5, 5, 0]))
    with pytest.raises(ValueError):
        p.plan(np.array([1, 1, 16]), np.array([5, 5, 0]))
    with pytest.raises(ValueError):
        p.plan(np.array([1, 40, 0]), np.array([5, 5, 0]))
    with pytest.raises(ValueError):
        R.RRTDubins(og, 50, 0.0)
    with pytest.raises(NotImplementedError):
        R.RRTDubins(og, 50, 2.0, costfn=lambda *a: 0.0)


def test_radius_set_overflow_is_reported_not_hidden():
    og = np.zeros((64, 64))
    p = R.RRTStar(og, 1500, 500.0, pbar=False, rewire="rrtstar")        # every vertex within the radius: > 1024 members
    with pytest.raises(MemoryError):
        p.plan(np.array([1, 1]), np.array([60, 60]))


def test_dubins_memo_table_equals_the_primitive():
    import torch
    L = _lib.lib()
    R_, nh, rho = 12, 8, 3.5
    nbytes = int(L.rrtk_dubins_table_bytes(R_, nh))
    N = (2 * R_ + 1) ** 2 * nh * nh
    assert nbytes >= N * 33
    tab = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
    _lib.check(L.rrtk_dubins_table_build(R_, nh, rho, tab.data_ptr(), torch.cuda.current_stream().cuda_stream), "table")
    torch.cuda.synchronize()
    raw = tab.cpu().numpy()
    tlen = raw[: 8 * N].view(np.float64)
    ttpq = raw[8 * N: 32 * N].view(np.float64).reshape(N, 3)
    tword = raw[32 * N: 33 * N]
    side = 2 * R_ + 1
    idx = np.arange(N)
    h1, h0, cell = idx % nh, (idx // nh) % nh, idx // (nh * nh)
    q = np.stack([np.zeros(N, int), np.zeros(N, int), h0, cell // side - R_, cell % side - R_, h1], axis=1)
    word, tpq, ln = O2.dubins(q, nh, rho)
    assert np.array_equal(tword.astype(np.int32), word)
    assert np.array_equal(tlen.view(np.int64), ln.view(np.int64)) and np.array_equal(ttpq.view(np.int64), tpq.view(np.int64))
    assert L.rrtk_dubins_table_bytes(0, nh) == 0 and L.rrtk_dubins_table_build(5, 300, 1.0, tab.data_ptr(), None) == -1


@pytest.mark.parametrize("threads,use_table", [(0, True), (128, True), (0, False)])
def test_device_batch2_matches_the_host_buffer_call(threads, use_table):
    W = H = 128
    n, nh, P = 500, 16, 5
    db = batch.DeviceBatch2("dubins", W, H, n, r_rewire=25.0, nheadings=nh, rho=4.0, ds=1.0, device=0, threads=threads, use_table=use_table)
    assert (db.table is not None) == use_table
    db.gen_worlds([worlds.world_seed(w) for w in range(P)])
    ogs = db.og.cpu().numpy()
    starts = np.array([[*np.argwhere(ogs[p] == 0)[7], p % nh] for p in range(P)])
    goals = np.array([[*np.argwhere(ogs[p] == 0)[-7], (5 * p) % nh] for p in range(P)])
    db.set_plans(batch.make_desc2(np.arange(P), starts, goals))
    db.seed_samples(np.arange(P)).seed_heads(50 + np.arange(P))
    res = db.run().download()
    smp = db.samples.cpu().numpy().astype(np.int64)
    hd = db.heads.cpu().numpy().astype(np.int64)
    for p in range(P):
        want = O2.plan("dubins", ogs[p], n, starts[p], goals[p], np.concatenate([smp[p], hd[p][:, None]], axis=1), star=True,
                       rewire=True, r_rewire=25.0, nh=nh, rho=4.0, ds=1.0)
        top = want["stats"]["j"] + want["stats"]["found"]
        assert res.stat("j")[p] == want["stats"]["j"] and res.stat("rewires")[p] == want["stats"]["rewires"]
        assert np.array_equal(res.parent[p, :top], want["parent"][:top])
        assert np.array_equal(res.cost[p, :top].view(np.int64), want["cost"][:top].view(np.int64))
    smem, blocks = db.footprint()
    assert smem > 0 and blocks >= 1
