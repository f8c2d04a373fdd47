"""GPU tests of the drop-in classes: same constructor / plan() calls as the reference's users make
(docs/getting-started.rst:50-72, tests/test_rrt.py:122-136), results compared with what the pinned
reference returned for the same seeds (tests/golden/seeded_api.npz)."""
import os

import networkx as nx
import numpy as np
import pytest

import rrtplanner_b200 as R
from oracle import rrt_oracle as O
from tests.conftest import GOLDEN, load_plan, golden_plans

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seeded():
    z = np.load(os.path.join(GOLDEN, "seeded_api.npz"))
    d = {k: z[k] for k in z.files}
    w, h = (int(v) for v in d["og_shape"])
    d["og"] = np.unpackbits(d["og"], axis=1)[:, :h].astype(np.int64)
    return d


def graph_matches(T, gv, z, prefix):
    assert int(gv) == int(z[f"{prefix}_gv"])
    assert [int(v) for v in T.nodes] == [int(v) for v in z[f"{prefix}_g_nodes"]]
    pts = np.stack([np.asarray(T.nodes[v]["pt"]) for v in T.nodes])
    assert pts.dtype == np.int64 and np.array_equal(pts, z[f"{prefix}_g_pts"])
    e = list(T.edges(data=True))
    assert [int(a) for a, _, _ in e] == [int(v) for v in z[f"{prefix}_g_eu"]]
    assert [int(b) for _, b, _ in e] == [int(v) for v in z[f"{prefix}_g_ev"]]
    assert np.array_equal(np.array([d["dist"] for _, _, d in e]), z[f"{prefix}_g_dist"])
    assert np.array_equal(np.array([float(d["cost"]) for _, _, d in e]), z[f"{prefix}_g_cost"])


@pytest.mark.parametrize("name", ["standard", "star", "informed"])
def test_seeded_public_api_matches_reference(seeded, name):
    og = seeded["og"]
    mk = {"standard": lambda: R.RRTStandard(og, 300, pbar=False, seed=42),
          "star": lambda: R.RRTStar(og, 300, 30, pbar=False, seed=43),
          "informed": lambda: R.RRTStarInformed(og, 500, 30, 10, pbar=False, seed=44)}[name]
    p = mk()
    for rep, (a, b) in enumerate(((seeded["xs"], seeded["xg"]), (seeded["xs2"], seeded["xg2"]))):
        T, gv = p.plan(a.copy(), b.copy())          # second call continues the generator (rrt.py:85)
        graph_matches(T, gv, seeded, f"{name}{rep}")
        path = p.route2gv(T, gv)
        assert path[0] == 0 and path[-1] == gv
        lines = p.vertices_as_ndarray(T, path)
        assert lines.shape == (len(path) - 1, 2, 2)
        if name == "informed":
            keys = sorted(p.ellipses)
            assert keys == [int(k) for k in seeded[f"{name}{rep}_ell_keys"]]
            vals = np.array([[p.ellipses[k][0][0], p.ellipses[k][0][1], p.ellipses[k][1], p.ellipses[k][2], p.ellipses[k][3]]
                             for k in keys]).reshape(-1, 5)
            assert np.allclose(vals, seeded[f"{name}{rep}_ell_vals"], rtol=1e-12, atol=1e-12)


def test_golden_streams_through_classes():
    """cfg1 (256^2, n=1000, r=50, seed 0): the class API reproduces the golden graphs."""
    for path in golden_plans():
        g = load_plan(path)
        if int(g["seed"]) < 0 or g["kind"] == "informed":
            continue
        og = g["og"]
        p = R.RRTStandard(og, g["n"], pbar=False, seed=int(g["seed"])) if g["kind"] == "standard" else \
            R.RRTStar(og, g["n"], float(g["r_rewire"]), pbar=False, seed=int(g["seed"]))
        T, gv = p.plan(g["xstart"], g["xgoal"])
        z = {f"x_{k}": v for k, v in g.items()}
        graph_matches(T, gv, z, "x")


def test_reference_test_suite_shapes():
    """The reference's own smoke tests (tests/test_rrt.py:97-136) through the drop-in."""
    rng = np.random.default_rng(0)
    for dtype in (int, float, np.uint32, np.float32):
        for topo in ("empty", "square"):
            og = np.zeros((43, 100), dtype=dtype)
            if topo == "square":
                og[43 // 4: 3 * 43 // 4, 100 // 4: 3 * 100 // 4] = 1
            base = R.RRT(og, 25)
            p1 = np.array([rng.integers(0, 43), rng.integers(0, 100)])
            p2 = np.array([rng.integers(0, 43), rng.integers(0, 100)])
            assert base.collisionfree(base.og, p1, p2) == O.collisionfree(og, p1, p2)
            pts = np.stack([rng.integers(0, 43, 10), rng.integers(0, 43, 10)], -1)
            assert np.array_equal(base.near(pts, p1), O.near_sorted(pts, p1))
            assert np.array_equal(base.within(pts, p1, 10), O.within(pts, p1, 10))
            for planner in (R.RRTStandard(og, 100, pbar=False), R.RRTStar(og, 100, r_rewire=50, pbar=False),
                            R.RRTStarInformed(og, 100, r_rewire=50, r_goal=5, pbar=False)):
                xs, xg = planner.sample_all_free(), planner.sample_all_free()
                T, gv = planner.plan(xs, xg)
                assert isinstance(T, nx.DiGraph) and T.number_of_nodes() in (100, 101)


def test_set_og_and_set_n_replan():
    og = R.perlin_occupancygrid(96, 96, seed=3)
    p = R.RRTStar(og, 200, 20, pbar=False, seed=1)
    xs, xg = R.worlds.start_goal(og, 1)
    T1, g1 = p.plan(xs, xg)
    og2 = np.zeros_like(og)
    p.set_og(og2)
    p.set_n(150)
    T2, g2 = p.plan(xs, xg)
    assert T2.number_of_nodes() == 151 and p.free.shape[0] == 96 * 96
    # equals a fresh reference-style run on the new grid with the generator where it now stands
    q = R.RRTStar(og2, 150, 20, pbar=False, seed=1)
    q.rand_gen.integers(0, np.argwhere(og == 0).shape[0], size=200)      # the draws plan #1 consumed
    T3, g3 = q.plan(xs, xg)
    assert g2 == g3 and list(T2.edges) == list(T3.edges)


def test_go2goal_standalone():
    og = R.perlin_occupancygrid(96, 96, seed=3)
    xs, xg = R.worlds.start_goal(og, 1)
    smp = O.sample_stream(og, 120, 5)
    t = O.plan_star(og, 120, 20, xs, xg, smp)
    from collections import defaultdict
    p = R.RRTStar(og, 120, 20, pbar=False)
    j = t.j
    pts = np.full((120, 2), np.iinfo(np.int64).min, dtype=np.int64)
    pts[:j] = t.points[:j]
    vc = np.full(120, np.inf)
    vc[:j] = t.vcosts[:j]
    vgoal, ch, par, P2, C2 = p.go2goal(vc, pts, xg, j, defaultdict(list), {})
    assert int(vgoal) == int(t.vgoal)
    if t.found:
        assert par[vgoal] == int(t.parents[t.vgoal]) and C2[vgoal] == t.vcosts[t.vgoal]


def test_subclass_injection_reproduces_the_golden_trees(golden_plan):
    """The reference's injection point (rrt.py:231; unitball :579 for the informed planner): drive the drop-in classes
    exactly the way tests/golden/make_golden.py drives the reference -- a subclass whose sample_all_free / unitball read
    pre-generated streams through one shared iteration counter -- and get the golden graph of every fixture."""
    g = golden_plan
    base = {"standard": R.RRTStandard, "star": R.RRTStar, "informed": R.RRTStarInformed}[g["kind"]]
    state = {"i": 0, "free": 0, "ball": 0}
    samples, balls = g["samples"], g["balls"]

    class Driven(base):
        def sample_all_free(self):
            i = state["i"]; state["i"] += 1; state["free"] += 1
            return samples[i].copy()

        def unitball(self):
            i = state["i"]; state["i"] += 1; state["ball"] += 1
            return balls[i].copy()

    if g["kind"] == "standard":
        p = Driven(g["og"], g["n"], pbar=False)
    elif g["kind"] == "star":
        p = Driven(g["og"], g["n"], float(g["r_rewire"]), pbar=False)
    else:
        p = Driven(g["og"], g["n"], float(g["r_rewire"]), float(g["r_goal"]), pbar=False)
    T, gv = p.plan(g["xstart"], g["xgoal"])
    graph_matches(T, gv, {f"x_{k}": v for k, v in g.items()}, "x")
    # the samplers were called as often as the reference's loop calls them: once per iteration, free space first
    assert state["i"] == g["n"]
    if g["kind"] == "informed":
        first = p.last_stats["first_solution_iter"]
        assert state["free"] == (g["n"] if first < 0 else first + 1) and state["ball"] == g["n"] - state["free"]
        assert sorted(p.ellipses) == [int(k) for k in g["ell_keys"]]
    else:
        assert state["ball"] == 0


def test_informed_planner_with_the_firing_rewire_matches_the_specification():
    """RRTStarInformed(..., rewire="rrtstar") through the class API == oracle/rewire_oracle.c driven with the draws the
    reference's loop would take from the planner's generator (free-space draws up to the first solution vertex, then two
    uniforms per iteration)."""
    from oracle import rewire_oracle as R2
    from oracle import rrt_oracle as O
    og = R.worlds.perlin_occupancygrid(120, 96, seed=33).astype(np.uint8)
    n, r, r_goal, seed = 1200, 25.0, 8.0, 5
    free = np.argwhere(og == 0)
    xs, xg = R.worlds.start_goal(og, 1)
    p = R.RRTStarInformed(og, n, r, r_goal, pbar=False, seed=seed, rewire="rrtstar")
    T, gv = p.plan(xs, xg)
    gen = np.random.default_rng(seed)
    probe_gen = np.random.default_rng(seed)
    smp = np.concatenate([free[probe_gen.integers(0, len(free), size=n)], np.zeros((n, 1), dtype=np.int64)], axis=1)
    rot = O.ellipse_rotation(xs, xg)
    kw = dict(star=True, rewire=True, r_rewire=r)
    first = R2.plan("euclid", og, n, [*xs, 0], [*xg, 0], smp, informed=dict(r_goal=r_goal, rot=rot, balls=None), **kw)["stats"]["first_solution_iter"]
    assert first >= 0
    gen.integers(0, len(free), size=first + 1)
    balls = np.zeros((n, 2))
    for i in range(first + 1, n):
        u = gen.uniform(0, 1)
        a = 2 * np.pi * gen.uniform(0, 1)
        balls[i] = [np.sqrt(u) * np.cos(a), np.sqrt(u) * np.sin(a)]
    want = R2.plan("euclid", og, n, [*xs, 0], [*xg, 0], smp, informed=dict(r_goal=r_goal, rot=rot, balls=balls), **kw)
    st = want["stats"]
    assert st["rewires"] > 10 and st["ell_iters"] == n - 1 - first
    assert p.last_stats["j"] == st["j"] and p.last_stats["rewires"] == st["rewires"] and p.last_stats["first_solution_iter"] == first
    j = st["j"]
    for v in range(1, j):
        assert np.array_equal(T.nodes[v]["pt"], want["pts"][v])
        (par, _, attr), = T.in_edges(v, data=True)
        assert par == want["parent"][v] and attr["cost"] == want["cost"][v]
    assert sorted(p.ellipses) == [int(k) for k in np.flatnonzero(~np.isnan(want["ell"]))]
    # the planner's generator was advanced exactly as the reference's loop would have advanced it
    assert p.rand_gen.uniform(0, 1) == gen.uniform(0, 1)


def test_static_collisionfree_sees_in_place_edits():
    """The static wrapper keeps the last grid on the device between calls, validated by content."""
    og = np.zeros((40, 60), dtype=np.int64)
    a, b = np.array([2, 3]), np.array([35, 50])
    assert R.RRT.collisionfree(og, a, b) is True
    assert R.RRT.collisionfree(og, a, b) is True          # served from the cached grid
    og[20, :] = 7                                           # in-place edit, same array object
    assert R.RRT.collisionfree(og, a, b) is False
    og[20, :] = 0
    assert R.RRT.collisionfree(og, a, b) is True
