"""CPU checks of oracle/rewire_oracle.c, the specification of the planners the reference advertises but does not
contain (rewire that fires; Dubins RRT / RRT*).  Parity status of those planners is UNPINNED (no reference code
exists); what can be pinned is pinned here: with the Euclidean model and rewire off the oracle must reproduce the
golden trees the real reference produced, and the Dubins primitive must satisfy its defining properties."""
import numpy as np
import pytest

from oracle import rewire_oracle as R
from oracle import rrt_oracle as O
from tests.conftest import golden_plans, load_plan

NH, RHO, DS = 16, 6.0, 1.0
SEG = [(0, 1, 0), (2, 1, 2), (0, 1, 2), (2, 1, 0), (2, 0, 2), (0, 2, 0)]


def test_deterministic_math_is_accurate():
    rng = np.random.default_rng(0)
    a, b = rng.uniform(-30, 30, 50000), rng.uniform(-30, 30, 50000)
    at2, sn, cs = R.math(a, b)
    assert np.abs(at2 - np.arctan2(a, b)).max() < 2e-14          # 16-term series: truncation < 2e-14
    assert np.abs(sn - np.sin(a)).max() < 1e-14 and np.abs(cs - np.cos(a)).max() < 1e-14
    at2, _, _ = R.math(np.array([0.0, 0.0, 1.0, -1.0, 0.0]), np.array([0.0, -1.0, 0.0, 0.0, 2.0]))
    assert np.allclose(at2, [0.0, np.pi, np.pi / 2, -np.pi / 2, 0.0], atol=1e-15)


def _advance(x, y, th, kind, ln, rho):
    if kind == 1:
        return x + rho * ln * np.cos(th), y + rho * ln * np.sin(th), th
    if kind == 0:
        return x + rho * (np.sin(th + ln) - np.sin(th)), y + rho * (np.cos(th) - np.cos(th + ln)), th + ln
    return x + rho * (np.sin(th) - np.sin(th - ln)), y + rho * (np.cos(th - ln) - np.cos(th)), th - ln


def test_every_dubins_word_reaches_the_target_and_the_shortest_is_chosen():
    rng = np.random.default_rng(1)
    seen = np.zeros(6, dtype=int)
    for _ in range(1500):
        q = np.array([rng.integers(0, 40), rng.integers(0, 40), rng.integers(0, NH),
                      rng.integers(0, 40), rng.integers(0, 40), rng.integers(0, NH)])
        ok, tpq = R.dubins_all(q, NH, RHO)
        assert ok[0] and ok[1]                                     # LSL and RSR always exist
        lens = np.full(6, np.inf)
        for w in np.flatnonzero(ok):
            x, y, th = float(q[0]), float(q[1]), q[2] * 2 * np.pi / NH
            for kind, ln in zip(SEG[w], tpq[w]):
                assert ln >= 0.0
                x, y, th = _advance(x, y, th, kind, ln, RHO)
            err = max(abs(x - q[3]), abs(y - q[4]), abs(np.angle(np.exp(1j * (th - q[5] * 2 * np.pi / NH)))))
            assert err < 1e-7, (q, R.WORDS[w], err)
            lens[w] = ((tpq[w][0] + tpq[w][1]) + tpq[w][2]) * RHO
            seen[w] += 1
        word, best, ln = R.dubins(q, NH, RHO)
        assert word[0] == int(np.argmin(lens)) and ln[0] == lens.min()
        assert ln[0] >= np.hypot(q[3] - q[0], q[4] - q[1]) - 1e-9  # never shorter than the straight line
    assert (seen > 300).all()


def test_dubins_points_are_continuous_and_end_on_target():
    rng = np.random.default_rng(2)
    for _ in range(200):
        q = np.array([rng.integers(0, 64), rng.integers(0, 64), rng.integers(0, NH),
                      rng.integers(0, 64), rng.integers(0, 64), rng.integers(0, NH)])
        _, _, ln = R.dubins(q, NH, RHO)
        s = np.linspace(0.0, ln[0], 400)
        p = R.dubins_points(q, NH, RHO, s)
        step = np.hypot(np.diff(p[:, 0]), np.diff(p[:, 1]))
        assert step.max() <= (s[1] - s[0]) * (1 + 1e-9) + 1e-12    # unit-speed curve: chord <= arc
        assert abs(p[0, 0] - q[0]) < 1e-12 and abs(p[0, 1] - q[1]) < 1e-12
        assert abs(p[-1, 0] - q[3]) < 1e-7 and abs(p[-1, 1] - q[4]) < 1e-7
        # curvature bound: heading changes by at most ds / rho
        dth = np.abs(np.diff(p[:, 2]))
        assert dth.max() <= (s[1] - s[0]) / RHO * (1 + 1e-9) + 1e-12


def test_dubins_free_agrees_with_a_python_walk():
    rng = np.random.default_rng(3)
    og = (rng.random((64, 64)) < 0.05).astype(np.uint8)
    q = np.stack([rng.integers(0, 64, 400), rng.integers(0, 64, 400), rng.integers(0, NH, 400),
                  rng.integers(0, 64, 400), rng.integers(0, 64, 400), rng.integers(0, NH, 400)], axis=1)
    got = R.dubins_free(og, q, NH, RHO, DS)
    _, _, ln = R.dubins(q, NH, RHO)
    for i in range(q.shape[0]):
        s = np.arange(int(np.floor(ln[i] / DS)) + 1) * DS
        p = R.dubins_points(q[i], NH, RHO, s)
        cx, cy = np.floor(p[:, 0] + 0.5).astype(int), np.floor(p[:, 1] + 0.5).astype(int)
        inside = (cx >= 0) & (cx < 64) & (cy >= 0) & (cy < 64)
        want = bool(inside.all() and not og[cx[inside], cy[inside]].any() and og[q[i, 3], q[i, 4]] == 0)
        assert got[i] == want
    assert 0.05 < got.mean() < 0.95


@pytest.mark.parametrize("path", [p for p in golden_plans() if "informed" not in p], ids=lambda p: p.split("plan_")[-1][:-4])
def test_euclid_without_rewire_reproduces_the_reference_trees(path):
    """The part of the specification the two models share is the reference's loop: pinned by the golden plans."""
    g = load_plan(path)
    n = g["n"]
    smp = np.concatenate([g["samples"], np.zeros((n, 1), dtype=np.int64)], axis=1)
    r = R.plan("euclid", g["og"], n, [*g["xstart"], 0], [*g["xgoal"], 0], smp, star=g["kind"] == "star", rewire=False,
               r_rewire=float(g["r_rewire"]))
    st = r["stats"]
    found = bool(st["found"])
    top = st["j"] + (1 if found else 0)
    assert int(g["vgoal"]) == (st["vgoal"] if found else 0)
    assert np.array_equal(r["pts"][:top], g["points"][:top])
    assert np.array_equal(r["cost"][:top].view(np.int64), g["vcosts"][:top].view(np.int64))
    assert np.array_equal(r["parent"][1:top], g["parents"][1:top])


@pytest.mark.parametrize("path", [p for p in golden_plans() if "informed" in p], ids=lambda p: p.split("plan_")[-1][:-4])
def test_informed_rule_without_rewire_reproduces_the_reference_trees(path):
    """The informed sampling rule of the specification (orc2_plan_informed) is the reference's: with the rewire off the
    RRTStarInformed trees the real reference produced come out bit for bit, ellipse records included."""
    g = load_plan(path)
    n = g["n"]
    smp = np.concatenate([g["samples"], np.zeros((n, 1), dtype=np.int64)], axis=1)
    inf = dict(r_goal=float(g["r_goal"]), rot=O.ellipse_rotation(g["xstart"], g["xgoal"]), balls=g["balls"])
    r = R.plan("euclid", g["og"], n, [*g["xstart"], 0], [*g["xgoal"], 0], smp, star=True, rewire=False, r_rewire=float(g["r_rewire"]),
               informed=inf)
    st = r["stats"]
    found = bool(st["found"])
    top = st["j"] + (1 if found else 0)
    assert int(g["vgoal"]) == (st["vgoal"] if found else 0)
    assert np.array_equal(r["pts"][:top], g["points"][:top])
    assert np.array_equal(r["cost"][:top].view(np.int64), g["vcosts"][:top].view(np.int64))
    assert np.array_equal(r["parent"][1:top], g["parents"][1:top])
    assert [int(k) for k in np.flatnonzero(~np.isnan(r["ell"]))] == [int(k) for k in g["ell_keys"]]
    # the probe stops at the first solution vertex and reports the iteration that accepted it
    probe = R.plan("euclid", g["og"], n, [*g["xstart"], 0], [*g["xgoal"], 0], smp, star=True, rewire=False,
                   r_rewire=float(g["r_rewire"]), informed=dict(inf, balls=None))
    first = st["first_solution_iter"]
    assert probe["stats"]["first_solution_iter"] == first
    if first >= 0:
        assert probe["stats"]["j"] <= st["j"] and st["ell_iters"] == n - 1 - first
        assert np.array_equal(probe["pts"][:probe["stats"]["j"]], r["pts"][:probe["stats"]["j"]])


@pytest.mark.parametrize("model", ["euclid", "dubins"])
def test_informed_rule_with_a_firing_rewire_keeps_the_tree_consistent(model):
    og = _world(21)
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(6)
    n = 900
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, NH, (n, 1))], axis=1)
    start, goal = [*free[10], 3], [*free[-10], 5]
    u, a = rng.uniform(0, 1, n), 2 * np.pi * rng.uniform(0, 1, n)
    balls = np.stack([np.sqrt(u) * np.cos(a), np.sqrt(u) * np.sin(a)], axis=1)
    inf = dict(r_goal=8.0, rot=O.ellipse_rotation(np.array(start[:2]), np.array(goal[:2])), balls=balls)
    kw = dict(star=True, r_rewire=20.0, nh=NH, rho=3.0, ds=DS)
    on = R.plan(model, og, n, start, goal, smp, rewire=True, informed=inf, **kw)
    _check_tree(on, og, model, NH, 3.0, DS)
    st = on["stats"]
    assert st["first_solution_iter"] >= 0 and st["ell_iters"] == n - 1 - st["first_solution_iter"] and st["rewires"] > 20
    ell = on["ell"]
    keys = np.flatnonzero(~np.isnan(ell))
    # the budget is the cost of a path through a solution vertex: never below the straight line.  (It is not monotone: the
    # reference picks the solution vertex by cost alone, rrt.py:627-633, and then adds that vertex's distance to the goal.)
    d = np.hypot(goal[0] - start[0], goal[1] - start[1])
    assert (ell[keys] >= d - 1e-9).all() and len(np.unique(ell[keys])) > 3


def _check_tree(r, og, model, nh, rho, ds):
    st = r["stats"]
    top = st["j"] + (1 if st["found"] else 0)
    pts, head, cost, elen, par = r["pts"], r["head"], r["cost"], r["elen"], r["parent"]
    assert par[0] == -1 and cost[0] == 0.0
    depth_ok = np.zeros(top, dtype=bool)
    depth_ok[0] = True
    for v in range(1, top):
        p = par[v]
        assert 0 <= p < st["j"] and p != v
        assert cost[v] == cost[p] + elen[v]                        # costs are consistent after every propagation
    for v in range(1, top):                                        # acyclic: every vertex reaches the root
        u, hops = v, 0
        while u != 0:
            u = par[u]
            hops += 1
            assert hops <= top
    q = np.stack([pts[par[1:top], 0], pts[par[1:top], 1], head[par[1:top]], pts[1:top, 0], pts[1:top, 1], head[1:top]], axis=1)
    if model == "dubins":
        _, _, ln = R.dubins(q, nh, rho)
        assert np.array_equal(ln, elen[1:top])
        assert R.dubins_free(og, q, nh, rho, ds).all()             # every tree edge is a free path parent -> child
    else:
        d = q[:, 3:5] - q[:, 0:2]
        assert np.array_equal(np.sqrt((d * d).sum(1).astype(np.float64)), elen[1:top])
        assert all(O.collisionfree(og, s[0:2], s[3:5]) for s in q.astype(np.int64))


def _world(seed, W=96, H=96):
    from rrtplanner_b200 import worlds
    return worlds.perlin_occupancygrid(W, H, seed=seed).astype(np.uint8)


@pytest.mark.parametrize("model", ["euclid", "dubins"])
def test_rewire_keeps_the_tree_consistent_and_lowers_costs(model):
    og = _world(7)
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(5)
    n = 600
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, NH, (n, 1))], axis=1)
    start, goal = [*free[10], 3], [*free[-10], 5]
    kw = dict(star=True, r_rewire=20.0, nh=NH, rho=3.0, ds=DS)
    off = R.plan(model, og, n, start, goal, smp, rewire=False, **kw)
    on = R.plan(model, og, n, start, goal, smp, rewire=True, **kw)
    _check_tree(off, og, model, NH, 3.0, DS)
    _check_tree(on, og, model, NH, 3.0, DS)
    assert off["stats"]["rewires"] == 0 and on["stats"]["rewires"] > 20
    assert on["stats"]["propagated"] > 0
    # the same vertices are accepted up to the first rewire; overall the rewired tree is cheaper
    j = min(on["stats"]["j"], off["stats"]["j"])
    assert np.nanmean(on["cost"][:j]) < np.nanmean(off["cost"][:j])


def test_dubins_rrt_without_star_uses_the_nearest_vertex():
    og = _world(9, 64, 64)
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(6)
    n = 200
    smp = np.concatenate([free[rng.integers(0, len(free), n)], rng.integers(0, NH, (n, 1))], axis=1)
    r = R.plan("dubins", og, n, [*free[0], 0], [*free[-1], 0], smp, star=False, rewire=False, nh=NH, rho=2.0, ds=DS)
    _check_tree(r, og, "dubins", NH, 2.0, DS)
    j = r["stats"]["j"]
    assert j > 20
    for v in range(1, j):
        d2 = ((r["pts"][:v] - r["pts"][v]) ** 2).sum(1)
        assert r["parent"][v] == int(np.argmin(d2))
