"""Pin the oracle (oracle/rrt_oracle.py and oracle/rrt_oracle.c) against the golden vectors the
real reference produced (tests/golden/make_golden.py) and against the reference's own
known-answer assertions (tests/test_rrt.py:68-71, 116-119).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle, rrt_oracle as O
from tests.conftest import GOLDEN


# ---- reference's own KATs ---------------------------------------------------------------
def test_within_unit_square_kat():
    pts = np.array([[0, 0], [1, 0], [1, 1], [0, 1]])
    assert O.within(pts, np.array([0.5, 0.5]), 1.0).shape[0] == 4      # tests/test_rrt.py:116-119


@pytest.mark.parametrize("dtype", [int, float, np.uint32, np.uint64, np.int32, np.int64, np.float32, np.float64])
def test_r2norm_kat(dtype):
    rng = np.random.default_rng(3)
    p = rng.integers(0, 10000, size=2).astype(dtype)
    assert np.isclose(O.r2norm(p), np.linalg.norm(p))                 # tests/test_rrt.py:68-71


def test_r2norm_bits_match_reference():
    z = np.load(os.path.join(GOLDEN, "misc.npz"))
    got = np.array([O.r2norm(v) for v in z["r2_in"]])
    assert (got.view(np.int64) == z["r2_out"].view(np.int64)).all()


def test_rotation_matches_reference():
    z = np.load(os.path.join(GOLDEN, "misc.npz"))
    for p, r in zip(z["rot_pairs"], z["rots"]):
        assert np.array_equal(O.ellipse_rotation(p[:2], p[2:]), r)


# ---- collision ---------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "collision_*.npz"))),
                         ids=lambda p: os.path.basename(p)[10:-4])
def test_collision_golden(path):
    z = np.load(path)
    og, segs, want = z["og"], z["segs"].astype(np.int64), z["free"]
    got_py = np.array([O.collisionfree(og, s[:2], s[2:]) for s in segs])
    assert (got_py == want).all()
    got_c, cells = c_oracle.collision_batch(og, segs)
    assert (got_c == want).all()
    # cells tested: python and C agree
    cells_py = np.array([O.first_hit(og, s[:2], s[2:])[1] for s in segs[:500]])
    assert (cells_py == cells[:500]).all()


def test_closed_form_cell_sequence():
    """k-th cell closed form == the walk, exhaustively for all segments in a 9x9 window and on
    random long segments."""
    big = np.zeros((2100, 2100), dtype=np.uint8)
    rng = np.random.default_rng(0)
    segs = [(4, 4, x, y) for x in range(9) for y in range(9)] + \
           [(x, y, 4, 4) for x in range(9) for y in range(9)]
    segs += [tuple(rng.integers(0, 2100, 4)) for _ in range(300)]
    for ax, ay, bx, by in segs:
        L = max(abs(bx - ax), abs(by - ay))
        for k in ({0, 1, L // 2, L - 1, L} if L > 40 else range(L + 1)):
            if k < 0:
                continue
            cx, cy = O.kth_cell((ax, ay), (bx, by), k)
            big[cx, cy] = 1
            ok, cells = O.first_hit(big, (ax, ay), (bx, by))
            big[cx, cy] = 0
            assert (not ok) and cells == k + 1, (ax, ay, bx, by, k)


# ---- queries -----------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["lattice", "sparse"])
def test_queries_golden(tag):
    z = np.load(os.path.join(GOLDEN, "queries.npz"))
    pts, qs, radii = z[f"{tag}_pts"], z[f"{tag}_qs"], z[f"{tag}_radii"]
    flat, lens, d2 = z[f"{tag}_within_flat"], z[f"{tag}_within_len"], z[f"{tag}_near0_d2"]
    off = 0
    for q, r, m, want_d2 in zip(qs, radii, lens, d2):
        want = flat[off:off + m]
        off += m
        assert np.array_equal(O.within(pts, q, r), want)
        assert np.array_equal(c_oracle.within(pts, len(pts), q, r), want)
        v = O.nearest(pts, q)
        vc, d2c = c_oracle.nearest(pts, len(pts), q)
        dd = pts[v] - q
        assert int(dd @ dd) == want_d2 == d2c
        assert v == vc == int(np.flatnonzero(((pts - q) ** 2).sum(1) == want_d2)[0])


def test_sampler_seed_stream():
    z = np.load(os.path.join(GOLDEN, "sampler_seed12345.npz"))
    assert np.array_equal(O.sample_stream(z["og"], 300, 12345), z["drawn"].astype(np.int64))


# ---- whole plans ---------------------------------------------------------------------------
def _run_py(g):
    k = g["kind"]
    if k == "standard":
        return O.plan_standard(g["og"], g["n"], g["xstart"], g["xgoal"], g["samples"])
    if k == "star":
        return O.plan_star(g["og"], g["n"], float(g["r_rewire"]), g["xstart"], g["xgoal"], g["samples"],
                           count_rewire=True)
    return O.plan_informed(g["og"], g["n"], float(g["r_rewire"]), float(g["r_goal"]), g["xstart"],
                           g["xgoal"], g["samples"], g["balls"])


def _run_c(g):
    rot = O.ellipse_rotation(g["xstart"], g["xgoal"]) if g["kind"] == "informed" else None
    return c_oracle.plan(g["kind"], g["og"], g["n"], g["xstart"], g["xgoal"], g["samples"],
                         float(g["r_rewire"]), float(g["r_goal"]),
                         g["balls"] if g["kind"] == "informed" else None, rot)


def _check_tree(t, g):
    assert t.points.shape[0] == int(g["rows"])
    assert np.array_equal(t.points, g["points"])
    assert np.array_equal(t.parents, g["parents"])
    assert np.array_equal(t.vcosts.view(np.int64), g["vcosts"].view(np.int64))   # bit-exact f64
    assert int(t.vgoal) == int(g["vgoal"]) == int(g["gv"])
    assert t.rewire_fired == 0
    # graph records of the public API (rrt.py:334-369)
    nodes, edges = O.graph_records(t)
    assert nodes == [int(v) for v in g["g_nodes"]]
    assert [e[0] for e in edges] == [int(v) for v in g["g_eu"]]
    assert [e[1] for e in edges] == [int(v) for v in g["g_ev"]]
    assert np.array_equal(np.array([e[2] for e in edges]), g["g_dist"])
    assert np.array_equal(np.array([e[3] for e in edges]), g["g_cost"])
    if g["kind"] == "informed":
        keys = sorted(t.ellipse_c)
        assert keys == [int(k) for k in g["ell_keys"]]


def test_plan_golden_python(golden_plan):
    _check_tree(_run_py(golden_plan), golden_plan)


def test_plan_golden_c(golden_plan):
    _check_tree(_run_c(golden_plan), golden_plan)


def test_c_and_python_agree_on_stats(golden_plan):
    a, b = _run_py(golden_plan), _run_c(golden_plan)
    assert (a.checks, a.cells, a.j) == (b.checks, b.cells, b.j)
    assert a.first_solution_iter == b.first_solution_iter and a.ellipse_iters == b.ellipse_iters
    for k in a.ellipse_c:
        assert a.ellipse_c[k] == b.ellipse_c[k]
