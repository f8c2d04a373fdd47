"""GPU parity of K8 (csrc/plan_rewire.cu, csrc/dubins.cuh) through the C ABI against oracle/rewire_oracle.c.

The planners tested here (rewire that fires, Dubins RRT / RRT*) do not exist in the reference, so the oracle is a
specification, not a restatement (parity UNPINNED).  The pinned part: Euclidean model with rewire off must give the
golden trees the real reference produced."""
import numpy as np
import pytest

from oracle import rewire_oracle as R
from rrtplanner_b200 import _lib, worlds
from tests.conftest import golden_plans, load_plan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context()
    yield c
    c.close()


def _queries(rng, nq, W, H, nh):
    return np.stack([rng.integers(0, W, nq), rng.integers(0, H, nq), rng.integers(0, nh, nq),
                     rng.integers(0, W, nq), rng.integers(0, H, nq), rng.integers(0, nh, nq)], axis=1).astype(np.int32)


@pytest.mark.parametrize("nh,rho", [(16, 6.0), (32, 2.5), (1, 10.0), (255, 0.75), (8, 40.0)])
def test_dubins_paths_bit_exact(ctx, nh, rho):
    rng = np.random.default_rng(nh)
    q = _queries(rng, 20000, 200, 160, nh)
    q[:50, 3:5] = q[:50, 0:2]                        # coincident positions
    q[50:100, 2] = q[50:100, 5]                      # equal headings
    q[100:150, 3] = q[100:150, 0]                    # vertical displacement
    word, tpq, ln = ctx.dubins_paths(q, nh, rho)
    w2, t2, l2 = R.dubins(q, nh, rho)
    assert np.array_equal(word, w2)
    assert np.array_equal(tpq.view(np.int64), t2.view(np.int64))
    assert np.array_equal(ln.view(np.int64), l2.view(np.int64))
    assert len(set(word.tolist())) >= (4 if nh > 1 else 2)


@pytest.mark.parametrize("nh,rho,ds", [(16, 6.0, 1.0), (32, 3.0, 0.5), (12, 12.0, 2.0)])
def test_dubins_collision_bit_exact(ctx, nh, rho, ds):
    og = worlds.perlin_occupancygrid(160, 130, seed=4).astype(np.uint8)
    ctx.set_grids(og[None])
    rng = np.random.default_rng(1)
    q = _queries(rng, 6000, 160, 130, nh)
    got = ctx.dubins_collision(q, nh, rho, ds)
    want = R.dubins_free(og, q, nh, rho, ds)
    assert np.array_equal(got, want)
    assert 0.02 < got.mean() < 0.98


def test_dubins_sample_bit_exact(ctx):
    nh, rho, ds = 16, 5.0, 0.75
    rng = np.random.default_rng(2)
    q = _queries(rng, 64, 90, 90, nh)
    xyth, cnt = ctx.dubins_sample(q, nh, rho, ds, 400)
    _, _, ln = R.dubins(q, nh, rho)
    for i in range(q.shape[0]):
        k = int(np.floor(ln[i] / ds)) + 1
        assert cnt[i] == k
        want = R.dubins_points(q[i], nh, rho, np.arange(min(k, 400)) * ds)
        assert np.array_equal(xyth[i, : min(k, 400)].view(np.int64), want.view(np.int64))


def _desc(start, goal):
    d = np.zeros(1, dtype=_lib.PLAN_DESC)
    d["start_x"], d["start_y"], d["goal_x"], d["goal_y"] = int(start[0]), int(start[1]), int(goal[0]), int(goal[1])
    d["reserved"][0, 0], d["reserved"][0, 1] = int(start[2]), int(goal[2])
    return d


def _run(ctx, model, og, n, start, goal, smp, star, rewire, r, nh=16, rho=1.0, ds=1.0):
    ctx.set_grids((og != 0).astype(np.uint8)[None])
    cfg = _lib.plan2_cfg(_lib.MODEL_DUBINS if model == "dubins" else _lib.MODEL_EUCLID, star, rewire, r, nh, rho, ds)
    pts, head, cost, elen, par, st = ctx.plan2(cfg, _desc(start, goal), n, samples=smp[None, :, :2].astype(np.int16),
                                               heads=smp[None, :, 2].astype(np.uint8))
    return pts[0], head[0], cost[0], elen[0], par[0], dict(zip(_lib.STAT2_NAMES, (int(v) for v in st[0])))


def _compare(got, want, model):
    pts, head, cost, elen, par, st = got
    ws = want["stats"]
    for k in ("j", "vgoal", "found", "accepted", "rewires", "propagated", "ring_members"):
        assert st[k] == ws[k], (k, st[k], ws[k])
    assert st["overflow"] == 0
    top = ws["j"] + (1 if ws["found"] else 0)
    assert np.array_equal(pts[:top].astype(np.int32), want["pts"][:top])
    assert np.array_equal(par[:top], want["parent"][:top])
    assert np.array_equal(cost[:top].view(np.int64), want["cost"][:top].view(np.int64))
    assert np.array_equal(elen[:top].view(np.int64), want["elen"][:top].view(np.int64))
    if model == "dubins":
        assert np.array_equal(head[:top].astype(np.int32), want["head"][:top])
    assert (pts[top:] == -32768).all() and (par[top:] == -1).all() and np.isinf(cost[top:]).all()


@pytest.mark.parametrize("path", [p for p in golden_plans() if "informed" not in p], ids=lambda p: p.split("plan_")[-1][:-4])
def test_euclid_without_rewire_reproduces_the_reference_trees(ctx, path):
    g = load_plan(path)
    n = g["n"]
    smp = np.concatenate([g["samples"], np.zeros((n, 1), dtype=np.int64)], axis=1)
    pts, head, cost, elen, par, st = _run(ctx, "euclid", g["og"], n, [*g["xstart"], 0], [*g["xgoal"], 0], smp,
                                          g["kind"] == "star", False, float(g["r_rewire"]))
    found = bool(st["found"])
    top = st["j"] + (1 if found else 0)
    assert int(g["vgoal"]) == (st["vgoal"] if found else 0)
    assert np.array_equal(pts[:top].astype(np.int64), g["points"][:top])
    assert np.array_equal(cost[:top].view(np.int64), g["vcosts"][:top].view(np.int64))
    assert np.array_equal(par[1:top].astype(np.int64), g["parents"][1:top])


def _run_informed(ctx, model, og, n, start, goal, smp, rewire, r, r_goal, rot, balls, nh=16, rho=1.0, ds=1.0):
    ctx.set_grids((og != 0).astype(np.uint8)[None])
    cfg = _lib.plan2_cfg(_lib.MODEL_DUBINS if model == "dubins" else _lib.MODEL_EUCLID, True, rewire, r, nh, rho, ds, informed=True, r_goal=r_goal)
    d = _desc(start, goal)
    d["rot"][0] = np.asarray(rot, dtype=np.float64).reshape(4)
    pts, head, cost, elen, par, st, ell = ctx.plan2(cfg, d, n, samples=smp[None, :, :2].astype(np.int16), heads=smp[None, :, 2].astype(np.uint8),
                                                    balls=None if balls is None else balls[None])
    return (pts[0], head[0], cost[0], elen[0], par[0], dict(zip(_lib.STAT2_NAMES, (int(v) for v in st[0])))), ell[0]


@pytest.mark.parametrize("path", [p for p in golden_plans() if "informed" in p], ids=lambda p: p.split("plan_")[-1][:-4])
def test_informed_without_rewire_reproduces_the_reference_trees(ctx, path):
    """K8 with the informed sampling rule and the rewire off == the RRTStarInformed trees of the real reference (pinned)."""
    from oracle import rrt_oracle as O
    g = load_plan(path)
    n = g["n"]
    smp = np.concatenate([g["samples"], np.zeros((n, 1), dtype=np.int64)], axis=1)
    rot = O.ellipse_rotation(g["xstart"], g["xgoal"])
    (pts, head, cost, elen, par, st), ell = _run_informed(ctx, "euclid", g["og"], n, [*g["xstart"], 0], [*g["xgoal"], 0], smp, False,
                                                          float(g["r_rewire"]), float(g["r_goal"]), rot, g["balls"])
    found = bool(st["found"])
    top = st["j"] + (1 if found else 0)
    assert int(g["vgoal"]) == (st["vgoal"] if found else 0)
    assert np.array_equal(pts[:top].astype(np.int64), g["points"][:top])
    assert np.array_equal(cost[:top].view(np.int64), g["vcosts"][:top].view(np.int64))
    assert np.array_equal(par[1:top].astype(np.int64), g["parents"][1:top])
    assert [int(k) for k in np.flatnonzero(~np.isnan(ell))] == [int(k) for k in g["ell_keys"]]


INFORMED_CASES = [
    # model, W, H, n, r, r_goal, rewire, nh, rho, world seed
    ("euclid", 96, 96, 900, 20.0, 8.0, True, 1, 1.0, 21),
    ("euclid", 200, 160, 2500, 30.0, 5.0, True, 1, 1.0, 22),
    ("euclid", 128, 128, 1500, 25.0, 40.0, True, 1, 1.0, 23),          # a goal region of ~5000 cells: a long solution list
    ("dubins", 96, 96, 700, 20.0, 8.0, True, 16, 3.0, 21),
    ("dubins", 128, 128, 1200, 30.0, 6.0, False, 16, 4.0, 24),
]


@pytest.mark.parametrize("case", INFORMED_CASES, ids=lambda c: f"{c[0]}_{c[1]}x{c[2]}_n{c[3]}_rg{c[5]:g}_rw{int(c[6])}")
def test_informed_plans_bit_exact_against_the_specification(ctx, case):
    from oracle import rrt_oracle as O
    model, W, H, n, r, r_goal, rewire, nh, rho, seed = case
    og = worlds.perlin_occupancygrid(W, H, seed=seed).astype(np.uint8)
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(seed)
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, nh, (n, 1))], axis=1)
    start, goal = [*free[7], 1 % nh], [*free[-7], 3 % nh]
    u, a = rng.uniform(0, 1, n), 2 * np.pi * rng.uniform(0, 1, n)
    balls = np.stack([np.sqrt(u) * np.cos(a), np.sqrt(u) * np.sin(a)], axis=1)
    rot = O.ellipse_rotation(np.array(start[:2]), np.array(goal[:2]))
    inf = dict(r_goal=r_goal, rot=rot, balls=balls)
    want = R.plan(model, og, n, start, goal, smp, star=True, rewire=rewire, r_rewire=r, nh=nh, rho=rho, ds=1.0, informed=inf)
    got, ell = _run_informed(ctx, model, og, n, start, goal, smp, rewire, r, r_goal, rot, balls, nh, rho, 1.0)
    _compare(got, want, model)
    assert got[5]["ell_iters"] == want["stats"]["ell_iters"] and got[5]["first_solution_iter"] == want["stats"]["first_solution_iter"]
    assert want["stats"]["first_solution_iter"] >= 0 and want["stats"]["ell_iters"] > n // 4
    assert np.array_equal(np.isnan(ell), np.isnan(want["ell"]))
    k = ~np.isnan(ell)
    assert np.array_equal(ell[k].view(np.int64), want["ell"][k].view(np.int64))
    # probe: stops at the first solution vertex, same prefix of the tree
    probe, _ = _run_informed(ctx, model, og, n, start, goal, smp, rewire, r, r_goal, rot, None, nh, rho, 1.0)
    wp = R.plan(model, og, n, start, goal, smp, star=True, rewire=rewire, r_rewire=r, nh=nh, rho=rho, ds=1.0, informed=dict(inf, balls=None))
    _compare(probe, wp, model)
    assert probe[5]["first_solution_iter"] == want["stats"]["first_solution_iter"] and probe[5]["ell_iters"] == 0


CASES = [
    # model, W, H, n, r, star, rewire, nh, rho, ds, world seed
    ("euclid", 96, 96, 600, 20.0, True, True, 1, 1.0, 1.0, 7),
    ("euclid", 128, 100, 1500, 30.0, True, True, 1, 1.0, 1.0, 8),
    ("euclid", 256, 256, 3000, 50.0, True, True, 1, 1.0, 1.0, 9),
    ("dubins", 96, 96, 500, 20.0, True, True, 16, 3.0, 1.0, 7),
    ("dubins", 96, 96, 500, 20.0, True, False, 16, 3.0, 1.0, 7),
    ("dubins", 96, 96, 300, 0.0, False, False, 16, 3.0, 1.0, 7),
    ("dubins", 128, 160, 1200, 30.0, True, True, 32, 5.0, 0.5, 10),
    ("dubins", 256, 256, 2500, 50.0, True, True, 16, 6.0, 1.0, 11),
    ("dubins", 64, 64, 400, 200.0, True, True, 8, 2.0, 1.0, 12),      # radius covers the grid: every vertex is a member
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}_{c[1]}x{c[2]}_n{c[3]}_r{c[4]:g}_s{int(c[5])}{int(c[6])}_h{c[7]}")
def test_plans_bit_exact_against_the_specification(ctx, case):
    model, W, H, n, r, star, rewire, nh, rho, ds, seed = case
    og = worlds.perlin_occupancygrid(W, H, seed=seed).astype(np.uint8)
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(seed)
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, nh, (n, 1))], axis=1)
    smp[5] = smp[2]                                                      # a duplicate sample
    start = [*free[rng.integers(0, len(free))], int(rng.integers(0, nh))]
    goal = [*free[rng.integers(0, len(free))], int(rng.integers(0, nh))]
    smp[9, :2] = start[:2]                                               # a sample on the start cell
    want = R.plan(model, og, n, start, goal, smp, star=star, rewire=rewire, r_rewire=r, nh=nh, rho=rho, ds=ds)
    got = _run(ctx, model, og, n, start, goal, smp, star, rewire, r, nh, rho, ds)
    _compare(got, want, model)
    if rewire:
        assert want["stats"]["rewires"] > 0


def test_full_size_dubins_rrtstar_plan(ctx):
    """BASELINE cfg5 shape: 512 x 512 world, n = 5000, r = 50."""
    og = worlds.perlin_occupancygrid(512, 512, seed=1000).astype(np.uint8)
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(0)
    n, nh = 5000, 16
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, nh, (n, 1))], axis=1)
    start, goal = [*free[100], 2], [*free[-100], 9]
    want = R.plan("dubins", og, n, start, goal, smp, star=True, rewire=True, r_rewire=50.0, nh=nh, rho=6.0, ds=1.0)
    got = _run(ctx, "dubins", og, n, start, goal, smp, True, True, 50.0, nh, 6.0, 1.0)
    _compare(got, want, "dubins")
    assert want["stats"]["j"] > 1000


def test_several_plans_in_one_launch(ctx):
    W = H = 128
    n, nh, nplans = 400, 16, 6
    ogs = np.stack([worlds.perlin_occupancygrid(W, H, seed=20 + w).astype(np.uint8) for w in range(nplans)])
    ctx.set_grids(ogs)
    desc = np.zeros(nplans, dtype=_lib.PLAN_DESC)
    smp = np.zeros((nplans, n, 3), dtype=np.int64)
    starts, goals = [], []
    for p in range(nplans):
        free = np.argwhere(ogs[p] == 0)
        rng = np.random.default_rng(100 + p)
        smp[p] = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, nh, (n, 1))], axis=1)
        s, g = [*free[3], p % nh], [*free[-3], (3 * p) % nh]
        starts.append(s); goals.append(g)
        desc[p]["world"] = p
        desc[p]["start_x"], desc[p]["start_y"], desc[p]["goal_x"], desc[p]["goal_y"] = s[0], s[1], g[0], g[1]
        desc[p]["reserved"][0], desc[p]["reserved"][1] = s[2], g[2]
    cfg = _lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 25.0, nh, 4.0, 1.0)
    pts, head, cost, elen, par, st = ctx.plan2(cfg, desc, n, samples=smp[:, :, :2].astype(np.int16), heads=smp[:, :, 2].astype(np.uint8))
    for p in range(nplans):
        want = R.plan("dubins", ogs[p], n, starts[p], goals[p], smp[p], star=True, rewire=True, r_rewire=25.0, nh=nh, rho=4.0, ds=1.0)
        _compare((pts[p], head[p], cost[p], elen[p], par[p], dict(zip(_lib.STAT2_NAMES, (int(v) for v in st[p])))), want, "dubins")


@pytest.mark.parametrize("bits", [False, True])
def test_pipelined_worlds_call_equals_the_resident_call(ctx, bits):
    """rrtk_ctx_plan2_worlds (worlds + plans in, trees + path records out, several chunks) == rrtk_ctx_set_grids + rrtk_ctx_plan2."""
    W = H = 96
    n, nh, nplans = 300, 8, 11
    ogs = np.stack([worlds.perlin_occupancygrid(W, H, seed=40 + w).astype(np.uint8) for w in range(nplans)])
    desc = np.zeros(nplans, dtype=_lib.PLAN_DESC)
    smp = np.zeros((nplans, n, 2), dtype=np.int16)
    hd = np.zeros((nplans, n), dtype=np.uint8)
    for p in range(nplans):
        free = np.argwhere(ogs[p] == 0)
        rng = np.random.default_rng(300 + p)
        smp[p] = free[rng.integers(0, len(free), n)]
        hd[p] = rng.integers(0, nh, n)
        desc[p]["world"] = p
        desc[p]["start_x"], desc[p]["start_y"], desc[p]["goal_x"], desc[p]["goal_y"] = *free[2], *free[-2]
        desc[p]["reserved"][0], desc[p]["reserved"][1] = p % nh, (5 * p) % nh
    cfg = _lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 20.0, nh, 3.0, 1.0)
    ctx.set_grids(ogs)
    pts, head, cost, elen, par, st = ctx.plan2(cfg, desc, n, samples=smp, heads=hd)
    grids = _lib.pack_grids_host(ogs) if bits else ogs
    cap = 64
    got = ctx.plan2_worlds(cfg, grids, W, H, desc, n, samples=smp, heads=hd, bits=bits, trees=True, paths=True, path_cap=cap, chunk=4)
    assert np.array_equal(got["pts"], pts) and np.array_equal(got["head"], head) and np.array_equal(got["parent"], par)
    assert np.array_equal(got["cost"].view(np.int64), cost.view(np.int64)) and np.array_equal(got["elen"].view(np.int64), elen.view(np.int64))
    names = list(_lib.STAT2_NAMES)
    keep = [i for i, nm in enumerate(names) if nm != "checks"]       # edge tests saved by the goal connection's pruning depend on timing
    assert np.array_equal(got["stats"][:, keep], st[:, keep])
    for p in range(nplans):
        v = int(st[p, names.index("vgoal")])
        chain = [v]
        while chain[-1] > 0:
            chain.append(int(par[p, chain[-1]]))
        chain = chain[::-1]
        assert got["len"][p] == len(chain)
        if len(chain) <= cap:
            assert list(got["path"][p, :len(chain)]) == chain and (got["path"][p, len(chain):] == -1).all()
            assert np.array_equal(got["xy"][p, :len(chain)], pts[p, chain]) and np.array_equal(got["path_head"][p, :len(chain)], head[p, chain])
            assert (got["path_head"][p, len(chain):] == 255).all()
        assert got["path_cost"][p].view(np.int64) == cost[p, v].view(np.int64)
    # paths only, seeds instead of sample arrays, default chunking: same statistics as the resident seed-mode call
    from rrtplanner_b200 import batch
    states = batch.seed_states(np.arange(nplans) + 7)
    want = ctx.plan2(cfg, desc, n, states=states, heads=hd)
    got2 = ctx.plan2_worlds(cfg, grids, W, H, desc, n, states=states, heads=hd, bits=bits, trees=False, paths=True, path_cap=cap)
    assert np.array_equal(got2["stats"][:, keep], want[5][:, keep])


def test_bad_arguments_raise(ctx):
    og = np.zeros((32, 32), dtype=np.uint8)
    ctx.set_grids(og[None])
    smp = np.zeros((1, 10, 2), dtype=np.int16)
    with pytest.raises(ValueError):
        ctx.plan2(_lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 5.0, 16, 0.0, 1.0), _desc([1, 1, 0], [5, 5, 0]), 10, samples=smp)
    with pytest.raises(ValueError):
        ctx.plan2(_lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 5.0, 16, 2.0, 1.0), _desc([1, 1, 16], [5, 5, 0]), 10, samples=smp)
    with pytest.raises(ValueError):
        ctx.plan2(_lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 5.0, 4, 2.0, 1.0), _desc([1, 1, 0], [5, 5, 0]), 10, samples=smp,
                  heads=np.full((1, 10), 7, dtype=np.uint8))
    with pytest.raises(ValueError):
        ctx.dubins_collision(np.array([[0, 0, 0, 40, 3, 0]]), 16, 2.0, 1.0)


@pytest.mark.parametrize("model", ["dubins", "euclid"])
def test_full_size_batch_properties(model):
    """BASELINE cfg5 at full size (1024 plans, 512 x 512 worlds, n = 5000, r = 50): properties that need no oracle.
    Costs are consistent with the edge lengths after all rewires, the graph is a tree rooted at the start, every edge
    parent -> child is free when re-tested by the stand-alone kernels (rrtk_dubins_collision / rrtk_collision_segments),
    stored Dubins lengths equal rrtk_dubins_paths, and the first plans equal the specification bit for bit."""
    import torch
    from rrtplanner_b200 import batch
    W = H = 512
    n, nh, rho, ds, r, P = 5000, 16, 6.0, 1.0, 50.0, 1024
    db = batch.DeviceBatch2(model, W, H, n, r_rewire=r, nheadings=nh, rho=rho, ds=ds, device=0)
    db.gen_worlds([worlds.world_seed(p) for p in range(P)])
    nfree = db.nfree()
    rng = np.random.default_rng(3)
    # start / goal: two free cells per world, taken from the device sampler's own stream
    pair = batch.DeviceBatch("star", W, H, 2, device=0)
    pair.bits, pair.rowcum = db.bits, db.rowcum
    pair.set_plans(batch.make_desc(np.arange(P), np.zeros((P, 2)), np.zeros((P, 2))))
    pair.seed_samples(5000 + np.arange(P))
    sg = pair.samples.cpu().numpy().astype(np.int64)
    hs = rng.integers(0, nh, size=(P, 2))
    starts = np.concatenate([sg[:, 0], hs[:, :1]], axis=1)
    goals = np.concatenate([sg[:, 1], hs[:, 1:]], axis=1)
    db.set_plans(batch.make_desc2(np.arange(P), starts, goals))
    db.seed_samples(np.arange(P))
    if model == "dubins":
        db.seed_heads(9000 + np.arange(P))
    res = db.run().download()
    assert (res.stat("overflow") == 0).all() and (nfree > 0).all()
    j, found = res.stat("j"), res.stat("found")
    top = j + found
    rows = np.arange(n + 1)[None, :]
    live = (rows < top[:, None]) & (rows >= 1)
    par = np.where(live, res.parent, 0)
    pc = np.take_along_axis(res.cost, par, axis=1)
    assert np.array_equal((pc + res.elen)[live].view(np.int64), res.cost[live].view(np.int64))      # cost = parent's cost + edge length
    assert (res.parent[:, 0] == -1).all() and (res.cost[:, 0] == 0).all()
    assert ((res.parent[live] >= 0) & (res.parent[live] < np.broadcast_to(j[:, None], live.shape)[live])).all()
    depth = np.zeros_like(res.parent)                                                                # every vertex reaches the root
    cur = par.copy()
    for _ in range(n):
        nz = cur > 0
        if not nz.any():
            break
        depth += nz
        cur = np.where(nz, np.take_along_axis(par, cur, axis=1), 0)
    assert not (cur > 0).any()
    assert np.median(res.stat("rewires")) > 1000
    # every edge re-tested by the stand-alone kernels, parent -> child
    pi, vi = np.nonzero(live)
    ctx = _lib.Context()
    try:
        for lo in range(0, P, 128):
            sel = (pi >= lo) & (pi < lo + 128)
            p_, v_ = pi[sel], vi[sel]
            u_ = res.parent[p_, v_]
            ctx.set_grids(db.og[lo: lo + 128].cpu().numpy())
            for q in range(lo, min(lo + 128, P)):                          # the host-buffer forms take one world per call
                m = p_ == q
                if model == "dubins":
                    qq = np.stack([res.pts[q, u_[m], 0], res.pts[q, u_[m], 1], res.head[q, u_[m]], res.pts[q, v_[m], 0],
                                   res.pts[q, v_[m], 1], res.head[q, v_[m]]], axis=1).astype(np.int32)
                    assert ctx.dubins_collision(qq, nh, rho, ds, world=q - lo).all()
                    if q % 64 == 0:
                        assert np.array_equal(ctx.dubins_paths(qq, nh, rho)[2].view(np.int64), res.elen[q, v_[m]].view(np.int64))
                else:
                    seg = np.concatenate([res.pts[q, u_[m]], res.pts[q, v_[m]]], axis=1).astype(np.int32)
                    assert ctx.collision(seg, world=q - lo).all()
    finally:
        ctx.close()
    smp = db.samples.cpu().numpy().astype(np.int64)
    hd = db.heads.cpu().numpy().astype(np.int64) if model == "dubins" else np.zeros((P, n), dtype=np.int64)
    ogs = db.og[:3].cpu().numpy()
    for p in range(3):
        want = R.plan(model, ogs[p], n, starts[p], goals[p], np.concatenate([smp[p], hd[p][:, None]], axis=1), star=True, rewire=True,
                      r_rewire=r, nh=nh, rho=rho, ds=ds)
        _compare((res.pts[p], res.head[p], res.cost[p], res.elen[p], res.parent[p],
                  dict(zip(_lib.STAT2_NAMES, (int(v) for v in res.stats[p])))), want, model)


@pytest.mark.parametrize("seed", range(24))
def test_randomised_sweep_against_the_specification(ctx, seed):
    """Random shapes, radii, turning radii, heading counts, steps and flag combinations (seeded): every result bit-exact."""
    rng = np.random.default_rng(1000 + seed)
    model = "dubins" if seed % 3 else "euclid"
    W, H = int(rng.integers(40, 200)), int(rng.integers(40, 200))
    n = int(rng.integers(50, 900))
    star = bool(rng.integers(0, 4))                       # mostly RRT*
    rewire = star and bool(rng.integers(0, 3))
    r = float(rng.uniform(5.0, 60.0)) if star else 0.0
    nh = int(rng.choice([1, 4, 16, 64])) if model == "dubins" else 1
    rho = float(rng.uniform(0.5, 12.0))
    ds = float(rng.choice([0.5, 1.0, 2.0]))
    og = worlds.perlin_occupancygrid(W, H, seed=int(rng.integers(0, 10000))).astype(np.uint8)
    free = np.argwhere(og == 0)
    if len(free) < 10:
        pytest.skip("world without free space")
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, nh, (n, 1))], axis=1)
    start = [*free[rng.integers(0, len(free))], int(rng.integers(0, nh))]
    goal = [*free[rng.integers(0, len(free))], int(rng.integers(0, nh))]
    want = R.plan(model, og, n, start, goal, smp, star=star, rewire=rewire, r_rewire=r, nh=nh, rho=rho, ds=ds)
    if want["stats"]["ring_members"] and r * r * np.pi > 1024 and n > 1024:
        pytest.skip("radius set may exceed the kernel's list")
    got = _run(ctx, model, og, n, start, goal, smp, star, rewire, r, nh, rho, ds)
    _compare(got, want, model)
