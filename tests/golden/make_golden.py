"""Generate the golden vectors under tests/golden/ by RUNNING THE REAL REFERENCE.

Run here (the build container), never on the GPU box:

    python tests/golden/make_golden.py

It imports ``/root/reference/rrtplanner/rrt.py`` by file path (the package import needs
matplotlib / pyfastnoisesimd) and records, as small compressed ``.npz`` files:

* ``collision_*.npz``   verdicts of the UNMODIFIED ``RRT.collisionfree`` (rrt.py:183-229)
* ``queries_*.npz``     results of the UNMODIFIED ``RRT.within`` (rrt.py:157-181) and the distance
                        ordering of ``RRT.near`` (rrt.py:131-155)
* ``plan_*.npz``        full trees + networkx graph records of RRTStandard / RRTStar /
                        RRTStarInformed ``plan()`` (rrt.py:386-447, 466-556, 653-758) on fixed sample
                        streams, with the two implementation-defined sort calls pinned to
                        ``kind="stable"``; nothing else is changed.

How the pin is applied without touching reference code: the reference module looks ``np`` up in
its globals at call time, so the loaded module's ``np`` is replaced by a proxy whose ``argsort``
forces ``kind="stable"`` and which forwards every other attribute to numpy.  Sample streams are
injected by subclassing ``sample_all_free`` / ``unitball`` (the reference's own override points).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from rrtplanner_b200 import worlds  # noqa: E402

REF = "/root/reference/rrtplanner/rrt.py"
warnings.filterwarnings("ignore", category=RuntimeWarning)  # int(inf) cast at rrt.py:408


class _StableNumpy:
    """numpy with argsort pinned to a stable sort; everything else forwarded."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def argsort(a, *args, **kw):
        return np.argsort(a, kind="stable")


def load_reference(pinned: bool):
    spec = importlib.util.spec_from_file_location("ref_rrt_pinned" if pinned else "ref_rrt", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if pinned:
        mod.np = _StableNumpy()
    return mod


def blobs(w, h, seed, density=0.25, dtype=np.uint8):
    """small deterministic blob world (independent of worlds.py so both generators get used)"""
    rng = np.random.default_rng(seed)
    og = np.zeros((w, h), dtype=np.int64)
    for _ in range(max(1, int(density * w * h / 60))):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        rx, ry = rng.integers(2, 7), rng.integers(2, 7)
        og[max(0, cx - rx):cx + rx, max(0, cy - ry):cy + ry] = 1
    return og.astype(dtype)


# ------------------------------------------------------------------------------------
def gen_collision(ref):
    cases = {
        "blobs_61x47": blobs(61, 47, 1),
        "perlin_128x96": worlds.perlin_occupancygrid(128, 96, seed=5).astype(np.float32) * 2.5,
        "square_43x100": None,
    }
    sq = np.zeros((43, 100), dtype=np.uint64)
    sq[43 // 4: 3 * 43 // 4, 100 // 4: 3 * 100 // 4] = 1      # tests/test_rrt.py:35-36
    cases["square_43x100"] = sq
    for ci, (name, og) in enumerate(cases.items()):
        w, h = og.shape
        rng = np.random.default_rng(ci + 7)
        segs = np.stack([rng.integers(0, w, 6000), rng.integers(0, h, 6000),
                         rng.integers(0, w, 6000), rng.integers(0, h, 6000)], axis=1)
        # short segments around a few anchor points: every direction/octant incl. degenerate
        local = []
        for ax, ay in [(w // 2, h // 2), (3, 3), (w - 4, h - 4)]:
            for dx in range(-3, 4):
                for dy in range(-3, 4):
                    local.append((ax, ay, ax + dx, ay + dy))
        segs = np.concatenate([segs, np.array(local)], axis=0).astype(np.int64)
        verdict = np.array([ref.RRT.collisionfree(og, s[:2].copy(), s[2:].copy()) for s in segs])
        np.savez_compressed(os.path.join(HERE, f"collision_{name}.npz"),
                            og=og, segs=segs.astype(np.int16), free=verdict)
        print("collision", name, og.dtype, verdict.mean())


def gen_queries(ref):
    rng = np.random.default_rng(11)
    out = {}
    for tag, (m, span) in {"lattice": (600, 40), "sparse": (900, 500)}.items():
        pts = rng.integers(0, span, size=(m, 2)).astype(np.int64)
        qs = rng.integers(0, span, size=(64, 2)).astype(np.int64)
        radii = rng.choice([3, 7.5, 10, 50], size=64)
        win = [ref.RRT.within(pts, q, r) for q, r in zip(qs, radii)]
        order0_d2 = []
        for q in qs:
            o = ref.RRT.near(pts, q)                     # unmodified: only d^2 of [0] is pinned
            d = pts[o[0]] - q
            order0_d2.append(int(d @ d))
        out[tag] = dict(pts=pts, qs=qs, radii=radii,
                        within_flat=np.concatenate(win), within_len=np.array([len(x) for x in win]),
                        near0_d2=np.array(order0_d2))
    # the reference's only known-answer vector, tests/test_rrt.py:116-119
    kat = ref.RRT.within(np.array([[0, 0], [1, 0], [1, 1], [0, 1]]), np.array([0.5, 0.5]), 1.0)
    assert kat.shape[0] == 4
    flat = {f"{t}_{k}": v for t, d in out.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(HERE, "queries.npz"), **flat)
    print("queries ok")


# ------------------------------------------------------------------------------------
def graph_arrays(T, gv):
    nodes = np.array(list(T.nodes), dtype=np.int64)
    pts = np.stack([np.asarray(T.nodes[v]["pt"], dtype=np.int64) for v in T.nodes])
    e = list(T.edges(data=True))
    eu = np.array([int(a) for a, _, _ in e], dtype=np.int64)
    ev = np.array([int(b) for _, b, _ in e], dtype=np.int64)
    ed = np.array([float(d["dist"]) for _, _, d in e])
    ec = np.array([float(d["cost"]) for _, _, d in e])
    return dict(g_nodes=nodes, g_pts=pts, g_eu=eu, g_ev=ev, g_dist=ed, g_cost=ec, gv=np.int64(gv))


def run_plan(refp, kind, og, n, xstart, xgoal, samples, balls=None, r=None, r_goal=None):
    """Run the pinned reference planner on explicit streams, capturing its local arrays by
    wrapping build_graph (rrt.py:334), which receives (vgoal, points, parents, vcosts)."""
    base = {"standard": refp.RRTStandard, "star": refp.RRTStar, "informed": refp.RRTStarInformed}[kind]
    state = {"i": 0}
    grabbed = {}

    class Driven(base):
        def sample_all_free(self):
            i = state["i"]; state["i"] += 1
            return samples[i].copy()

        def unitball(self):
            i = state["i"]; state["i"] += 1
            return balls[i].copy()

        def build_graph(self, vgoal, points, parents, vcosts):
            grabbed.update(vgoal=vgoal, points=points.copy(), parents=dict(parents),
                           vcosts=vcosts.copy())
            return super().build_graph(vgoal, points, parents, vcosts)

    if kind == "standard":
        p = Driven(og, n, pbar=False)
    elif kind == "star":
        p = Driven(og, n, r, pbar=False)
    else:
        p = Driven(og, n, r, r_goal, pbar=False)
    T, gv = p.plan(np.asarray(xstart, dtype=np.int64), np.asarray(xgoal, dtype=np.int64))
    rows = grabbed["points"].shape[0]
    par = np.full(rows, -1, dtype=np.int64)
    for c, q in grabbed["parents"].items():
        if q is not None:
            par[int(c)] = int(q)
    rec = dict(points=grabbed["points"], vcosts=grabbed["vcosts"], parents=par,
               vgoal=np.int64(grabbed["vgoal"]), rows=np.int64(rows))
    rec.update(graph_arrays(T, gv))
    if kind == "informed":
        keys = np.array(sorted(p.ellipses), dtype=np.int64)
        rec["ell_keys"] = keys
        rec["ell_vals"] = np.array([[p.ellipses[k][0][0], p.ellipses[k][0][1], p.ellipses[k][1],
                                     p.ellipses[k][2], p.ellipses[k][3]] for k in keys]).reshape(-1, 5)
    return rec


def save_plan(name, kind, og, n, xstart, xgoal, samples, rec, balls=None, r=0.0, r_goal=0.0, seed=-1):
    np.savez_compressed(
        os.path.join(HERE, f"plan_{name}.npz"),
        kind=kind, og=np.packbits(og != 0, axis=1), og_shape=np.array(og.shape), n=np.int64(n),
        xstart=np.asarray(xstart, dtype=np.int64), xgoal=np.asarray(xgoal, dtype=np.int64),
        samples=np.asarray(samples, dtype=np.int16), r_rewire=np.float64(r), r_goal=np.float64(r_goal),
        seed=np.int64(seed), balls=(np.zeros((0, 2)) if balls is None else balls), **rec)
    print("plan", name, kind, "n", n, "rows", int(rec["rows"]), "vgoal", int(rec["vgoal"]),
          "filled", int(np.sum(rec["points"][:, 0] != np.iinfo(np.int64).min)))


def free_list(og):
    return np.argwhere(og == 0)


def seeded_stream(og, n, seed):
    f = free_list(og)
    return f[np.random.default_rng(seed).integers(0, f.shape[0], size=n)]


def pick_pair(og, seed):
    f = free_list(og)
    rng = np.random.default_rng(seed)
    a, b = f[rng.integers(0, f.shape[0])], f[rng.integers(0, f.shape[0])]
    return a, b


def gen_plans(refp, ref_plain):
    # (1) the reference's own fixtures (tests/test_rrt.py:20-41,125-129): n=100, r=50, r_goal=5
    sq = np.zeros((43, 100), dtype=np.float64)
    sq[43 // 4: 3 * 43 // 4, 100 // 4: 3 * 100 // 4] = 1
    empty = np.zeros((100, 43), dtype=np.int32)
    for gname, og in (("square43x100", sq), ("empty100x43", empty)):
        xs, xg = pick_pair(og, 3)
        smp = seeded_stream(og, 100, 0)
        balls = unit_balls(100, 5)
        for kind in ("standard", "star", "informed"):
            rec = run_plan(refp, kind, og, 100, xs, xg, smp, balls, r=50, r_goal=5)
            save_plan(f"{gname}_{kind}_n100", kind, og, 100, xs, xg, smp, rec, balls, 50, 5, seed=0)

    # (2) seed equivalence: the UNTOUCHED sampler of the reference with seed=s draws exactly
    #     seeded_stream(og, n, s) (rrt.py:85,240) -- recorded so the oracle's sample_stream and the
    #     device PCG64 can be checked against it.
    og = blobs(96, 128, 4)
    p = ref_plain.RRT(og, 10, pbar=False, seed=12345)
    drawn = np.stack([p.sample_all_free() for _ in range(300)])
    assert (drawn == seeded_stream(og, 300, 12345)).all()
    np.savez_compressed(os.path.join(HERE, "sampler_seed12345.npz"), og=og, drawn=drawn.astype(np.int16))

    # (3) blob world, medium n, small radius
    xs, xg = pick_pair(og, 8)
    smp = seeded_stream(og, 400, 21)
    balls = unit_balls(400, 6)
    for kind, r, rg in (("standard", 0, 0), ("star", 25, 0), ("informed", 25, 9)):
        rec = run_plan(refp, kind, og, 400, xs, xg, smp, balls, r=r, r_goal=rg)
        save_plan(f"blobs96x128_{kind}_n400", kind, og, 400, xs, xg, smp, rec, balls, r, rg, seed=21)

    # (4) cfg1 shape: 256^2 synthetic world, n=1000, r=50 (BASELINE.json configs[0])
    og = worlds.perlin_occupancygrid(256, 256, seed=worlds.world_seed(0))
    xs, xg = worlds.start_goal(og, 0)
    smp = seeded_stream(og, 1000, 0)
    balls = unit_balls(1000, 7)
    for kind, r, rg in (("standard", 0, 0), ("star", 50, 0), ("informed", 50, 5)):
        rec = run_plan(refp, kind, og, 1000, xs, xg, smp, balls, r=r, r_goal=rg)
        save_plan(f"cfg1_256_{kind}_n1000", kind, og, 1000, xs, xg, smp, rec, balls, r, rg, seed=0)

    # (5) adversarial stream: duplicates, the start point itself, obstacle cells, the goal itself
    og = blobs(48, 40, 9)
    f = free_list(og)
    occ = np.argwhere(og != 0)
    rng = np.random.default_rng(77)
    xs, xg = f[5], f[-7]
    base = f[rng.integers(0, f.shape[0], size=90)]
    smp = np.concatenate([xs[None], xs[None], base[:30], base[:30], occ[rng.integers(0, occ.shape[0], 10)],
                          base[30:], xg[None], xg[None]], axis=0)
    n = smp.shape[0]
    balls = unit_balls(n, 8)
    for kind, r, rg in (("standard", 0, 0), ("star", 12, 0), ("informed", 12, 6)):
        rec = run_plan(refp, kind, og, n, xs, xg, smp[:n], balls, r=r, r_goal=rg)
        save_plan(f"adversarial48x40_{kind}_n{n}", kind, og, n, xs, xg, smp[:n], rec, balls, r, rg)

    # (6) wall world ("impossible", tests/test_rrt.py:38-41) with every slot filled: goal not
    #     connectable, reference returns vgoal = 0 and n rows (rrt.py:328-331).
    og = np.zeros((40, 30), dtype=np.int64)
    og[20:21] = 1
    left = np.argwhere(og[:20] == 0)
    rng = np.random.default_rng(5)
    smp = left[rng.permutation(left.shape[0])[:80]]
    xs, xg = np.array([3, 3]), np.array([35, 20])
    n = 40
    balls = unit_balls(n, 9)
    for kind, r, rg in (("standard", 0, 0), ("star", 10, 0)):
        rec = run_plan(refp, kind, og, n, xs, xg, smp[:n], balls, r=r, r_goal=rg)
        assert int(rec["vgoal"]) == 0 and int(rec["rows"]) == n
        save_plan(f"wall40x30_{kind}_n{n}", kind, og, n, xs, xg, smp[:n], rec, balls, r, rg)


def gen_seeded(refp):
    """The reference driven ONLY through its public constructor/plan API with a seed (no stream
    injection): what the drop-in classes must reproduce.  Two successive plan() calls on one object
    check that the generator state carries over (rrt.py:85)."""
    og = worlds.perlin_occupancygrid(160, 120, seed=worlds.world_seed(7))
    xs, xg = worlds.start_goal(og, 7)
    xs2, xg2 = worlds.start_goal(og, 8)
    out = {"og": np.packbits(og != 0, axis=1), "og_shape": np.array(og.shape), "xs": xs, "xg": xg, "xs2": xs2, "xg2": xg2}
    cases = {
        "standard": lambda: refp.RRTStandard(og, 300, pbar=False, seed=42),
        "star": lambda: refp.RRTStar(og, 300, 30, pbar=False, seed=43),
        "informed": lambda: refp.RRTStarInformed(og, 500, 30, 10, pbar=False, seed=44),
    }
    for name, mk in cases.items():
        p = mk()
        for rep, (a, b) in enumerate(((xs, xg), (xs2, xg2))):
            T, gv = p.plan(a.copy(), b.copy())
            for k, v in graph_arrays(T, gv).items():
                out[f"{name}{rep}_{k}"] = v
            if name == "informed":
                keys = np.array(sorted(p.ellipses), dtype=np.int64)
                out[f"{name}{rep}_ell_keys"] = keys
                out[f"{name}{rep}_ell_vals"] = np.array([[p.ellipses[k][0][0], p.ellipses[k][0][1], p.ellipses[k][1],
                                                         p.ellipses[k][2], p.ellipses[k][3]] for k in keys]).reshape(-1, 5)
            print("seeded", name, rep, "gv", int(gv), "nodes", T.number_of_nodes(), "edges", T.number_of_edges())
    np.savez_compressed(os.path.join(HERE, "seeded_api.npz"), **out)


def unit_balls(n, seed):
    """n unit-ball points produced exactly as RRTStarInformed.unitball does (rrt.py:579-587)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 2))
    for i in range(n):
        r = rng.uniform(0, 1)
        theta = 2 * np.pi * rng.uniform(0, 1)
        out[i] = (np.sqrt(r) * np.cos(theta), np.sqrt(r) * np.sin(theta))
    return out


def gen_misc(ref):
    # r2norm known answers (tests/test_rrt.py:68-71) on integer vectors: value bits
    rng = np.random.default_rng(2)
    v = rng.integers(-3000, 3000, size=(256, 2)).astype(np.int64)
    out = np.array([ref.r2norm(x) for x in v])
    assert np.allclose(out, np.linalg.norm(v, axis=1))
    # ellipse rotation (rrt.py:601-613) for a spread of start/goal pairs incl. axis-aligned ones
    pairs = rng.integers(0, 300, size=(64, 4)).astype(np.int64)
    pairs[0] = (5, 7, 90, 7); pairs[1] = (5, 7, 5, 70); pairs[2] = (50, 7, 5, 7); pairs[3] = (5, 70, 5, 7)
    pl = ref.RRTStarInformed(np.zeros((4, 4)), 4, 1, 1, pbar=False)
    rots = np.stack([pl.rotation_to_world_frame(p[:2], p[2:]) for p in pairs])
    np.savez_compressed(os.path.join(HERE, "misc.npz"), r2_in=v, r2_out=out, rot_pairs=pairs, rots=rots)
    print("misc ok")


if __name__ == "__main__":
    ref_plain = load_reference(pinned=False)
    ref_pinned = load_reference(pinned=True)
    gen_collision(ref_plain)
    gen_queries(ref_plain)
    gen_misc(ref_plain)
    gen_plans(ref_pinned, ref_plain)
    gen_seeded(ref_pinned)
