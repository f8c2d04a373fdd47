"""CPU oracle for the rrtplanner tree-expansion hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch *restatement* (numpy + Python loops, Numba for the two
functions the reference JITs) of the algorithms in the reference's ``rrtplanner/rrt.py``.
It is the checker the CUDA path is compared with; it is never imported by the product
package ``rrtplanner_b200`` (only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it).

Parity status: **pinned**.  ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference/rrtplanner/rrt.py`` (by file path), runs it on fixed sample streams and commits
the outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this restatement (and the C
restatement ``oracle/rrt_oracle.c``) against every one of those vectors and against the two
known-answer assertions the reference's own test-suite holds (``tests/test_rrt.py:68-71,116-119``).

Pinned conventions; every deviation from the literal reference is a
place where the reference itself is implementation-defined or undefined:

* nearest vertex  = lowest index among vertices at minimum distance.  The reference takes element
  0 of an *unstable* ``np.argsort`` (``rrt.py:150-155``), so the tie winner there depends on the
  numpy build; the golden vectors come from the reference with ``near`` switched to
  ``kind="stable"`` (same rule).
* goal connection = candidates in ascending (cost, index) order, unfilled slots skipped.  The
  reference walks an unstable ``np.argsort(costs)`` (``rrt.py:317``) and, if no filled vertex sees
  the goal, indexes the grid with the INT64_MIN sentinel (out-of-bounds read, ``rrt.py:318``).
  Here that case is defined as "goal not connected": ``vgoal = 0`` and no goal row is appended
  (this is what the reference returns when every slot is filled, ``rrt.py:328-331``).
* samples are an explicit input (the pre-generated stream); ``sample_stream`` reproduces what
  ``RRT.sample_all_free`` (``rrt.py:231-240``) would draw for a given seed.

The "rewire" block of the reference (``rrt.py:532-546`` / ``:732-742``) compares
``vcosts[vn] + d < vcosts[vn]`` with ``d >= 0`` and therefore never changes the tree with the
default cost function; it is restated as a counter of how often the predicate fires (always 0)
so the claim is checked rather than assumed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

try:  # the reference JITs r2norm and collisionfree (rrt.py:10, rrt.py:183-184); do the same
    import numba as _nb

    _njit = _nb.njit(cache=False)
    HAVE_NUMBA = True
except Exception:  # pragma: no cover - numba is part of the image
    HAVE_NUMBA = False

    def _njit(f):
        return f


UNFILLED = np.iinfo(np.int64).min  # what int(inf) becomes in the reference (rrt.py:81,408)


# --------------------------------------------------------------------------------------
# leaf primitives
# --------------------------------------------------------------------------------------
@_njit
def _norm2(dx, dy):
    # rrt.py:10-24 -- exact integer d^2, one correctly rounded f64 sqrt
    return math.sqrt(dx * dx + dy * dy)


def r2norm(v) -> float:
    """Euclidean length of a 2-vector (rrt.py:10-24)."""
    return float(_norm2(v[0], v[1]))


@_njit
def _walk(og, ax, ay, bx, by):
    """Integer line walk of rrt.py:183-229.  Returns k >= 0: index of the first occupied cell
    (cells are numbered 0..L along the walk), or -(L + 1) = -(cells tested) when the whole walk,
    end cell included, is free."""
    adx = abs(bx - ax)
    ady = abs(by - ay)
    stepx = 1 if ax < bx else -1
    stepy = 1 if ay < by else -1
    acc = adx - ady
    x = ax
    y = ay
    k = 0
    while True:
        if og[x, y] != 0:
            return k
        if x == bx and y == by:
            return -(k + 1)
        twice = 2 * acc
        if twice >= -ady:
            acc -= ady
            x += stepx
        if twice <= adx:
            acc += adx
            y += stepy
        k += 1


def first_hit(og: np.ndarray, a, b) -> Tuple[bool, int]:
    """(free, cells_tested).  ``free`` is the reference's verdict (rrt.py:183-229);
    ``cells_tested`` is how many grid cells the reference's loop read before returning
    (first occupied cell inclusive, or all L+1 cells)."""
    r = int(_walk(og, int(a[0]), int(a[1]), int(b[0]), int(b[1])))
    if r >= 0:
        return False, r + 1
    return True, -r


def collisionfree(og: np.ndarray, a, b) -> bool:
    """True iff no non-zero cell lies on the walk a -> b, both ends included (rrt.py:183-229)."""
    return _walk(og, int(a[0]), int(a[1]), int(b[0]), int(b[1])) < 0


def kth_cell(a, b, k: int) -> Tuple[int, int]:
    """Closed form of the k-th cell (k = 0..L, L = max(|dx|,|dy|)) visited by the walk of
    rrt.py:183-229: the major axis advances one cell per step, the minor axis offset is
    floor((2*k*minor + major) / (2*major)) (ties round away from the start).  Checked
    exhaustively against ``_walk`` in tests/test_oracle.py."""
    ax, ay, bx, by = int(a[0]), int(a[1]), int(b[0]), int(b[1])
    adx, ady = abs(bx - ax), abs(by - ay)
    sx = 1 if ax < bx else -1
    sy = 1 if ay < by else -1
    if adx >= ady:
        major, minor = adx, ady
        off = (2 * k * minor + major) // (2 * major) if major else 0
        return ax + sx * k, ay + sy * off
    major, minor = ady, adx
    off = (2 * k * minor + major) // (2 * major)
    return ax + sx * off, ay + sy * k


def near_sorted(points: np.ndarray, x: np.ndarray, stable: bool = True) -> np.ndarray:
    """All row indices ordered by distance to x (rrt.py:131-155).  ``stable=True`` is the pinned
    rule (lowest index first among equals); ``stable=False`` is the literal reference call."""
    delta = points - x
    dist = np.linalg.norm(delta, axis=1)
    return np.argsort(dist, kind="stable") if stable else np.argsort(dist)


def nearest(points: np.ndarray, x: np.ndarray) -> int:
    """Pinned nearest vertex: element 0 of the stable ordering (rrt.py:422,503,703)."""
    return int(near_sorted(points, x, stable=True)[0])


def within(points: np.ndarray, x: np.ndarray, r: float) -> np.ndarray:
    """Ascending indices with squared distance strictly below r*r (rrt.py:157-181)."""
    delta = points - x
    sq = delta[:, 0] * delta[:, 0] + delta[:, 1] * delta[:, 1]
    hit = np.flatnonzero(sq < r * r)
    return hit


def edge_cost(vcosts: np.ndarray, points: np.ndarray, v: int, x: np.ndarray) -> float:
    """Default cost function: cost-to-come of v plus straight-line length (rrt.py:70-78)."""
    return vcosts[v] + _norm2(points[v, 0] - x[0], points[v, 1] - x[1])


# --------------------------------------------------------------------------------------
# sampling (rrt.py:27-44, 64, 231-240)
# --------------------------------------------------------------------------------------
def free_cells(og: np.ndarray) -> np.ndarray:
    """Row-major list of free cells, ``np.argwhere(og == 0)`` (rrt.py:64)."""
    return np.argwhere(og == 0)


def sample_stream(og: np.ndarray, n: int, seed: int = 0) -> np.ndarray:
    """The n points ``RRT.sample_all_free`` returns over one plan() of a planner built with
    ``seed`` (rrt.py:85,240): n successive bounded draws from PCG64 index the free list."""
    cells = free_cells(og)
    idx = np.random.default_rng(seed).integers(0, cells.shape[0], size=n)
    return cells[idx]


# --------------------------------------------------------------------------------------
# result record
# --------------------------------------------------------------------------------------
@dataclass
class Tree:
    """Arrays of one finished plan, shaped like the reference's locals after go2goal.

    points  (rows, 2) int64   rows = n, or n + 1 when the goal was connected (rrt.py:320-323)
    vcosts  (rows,)   float64 inf in unfilled rows
    parents (rows,)   int64   -1 = no entry in the reference's ``parents`` dict, root has -1 too
    j       number of filled vertices before goal connection
    vgoal   goal vertex id (j when connected, else 0)
    """

    points: np.ndarray
    vcosts: np.ndarray
    parents: np.ndarray
    j: int
    vgoal: int
    found: bool
    n: int
    checks: int = 0          # collision checks the reference would issue (lazy order)
    cells: int = 0           # grid cells those checks read
    rewire_fired: int = 0    # times the (dead) rewire predicate was true
    first_solution_iter: int = -1
    ellipse_iters: int = 0
    ellipse_c: Dict[int, float] = field(default_factory=dict)  # j -> cbest used (rrt.py:701)
    accepted: Optional[np.ndarray] = None  # per-iteration accept flags (trace)

    def path(self):
        """Vertex ids root -> vgoal following parents (what nx.shortest_path yields on a tree)."""
        out = [int(self.vgoal)]
        while out[-1] != 0:
            p = int(self.parents[out[-1]])
            if p < 0:
                return None
            out.append(p)
        return out[::-1]

    def path_cost(self) -> float:
        return float(self.vcosts[self.vgoal])


def _fresh(n: int, xstart):
    pts = np.full((n, 2), UNFILLED, dtype=np.int64)          # rrt.py:408
    cst = np.full((n,), np.inf)                               # rrt.py:409
    par = np.full((n,), -1, dtype=np.int64)
    pts[0] = xstart
    cst[0] = 0.0
    return pts, cst, par


class _Counter:
    __slots__ = ("checks", "cells")

    def __init__(self):
        self.checks = 0
        self.cells = 0

    def free(self, og, a, b) -> bool:
        ok, c = first_hit(og, a, b)
        self.checks += 1
        self.cells += c
        return ok


def connect_goal(og, vcosts, points, parents, xgoal, j, cnt: Optional[_Counter] = None):
    """Goal connection, rrt.py:284-332, pinned (see module docstring).  Returns
    (vgoal, found, points, vcosts, parents) with the reference's row layout."""
    cnt = cnt or _Counter()
    n = points.shape[0]
    togo = np.empty(n)
    for i in range(n):                                         # rrt.py:313-314
        togo[i] = edge_cost(vcosts, points, i, xgoal) if i < j else np.inf
    order = np.argsort(togo, kind="stable")
    for idx in order:
        idx = int(idx)
        if idx >= j:
            break                                              # unfilled: defined as "stop"
        if cnt.free(og, points[idx], xgoal):
            vgoal = j
            points = np.concatenate((points, np.asarray(xgoal, dtype=np.int64)[None, :]), axis=0)
            vcosts = np.concatenate((vcosts, [togo[idx]]), axis=0)
            parents = np.concatenate((parents, [-1]), axis=0)
            points[vgoal] = xgoal
            vcosts[vgoal] = togo[idx]
            parents[vgoal] = idx
            return vgoal, True, points, vcosts, parents
    return 0, False, points, vcosts, parents


# --------------------------------------------------------------------------------------
# planners
# --------------------------------------------------------------------------------------
def plan_standard(og, n, xstart, xgoal, samples, trace=False) -> Tree:
    """RRTStandard.plan on an explicit sample stream (rrt.py:386-447)."""
    xstart = np.asarray(xstart, dtype=np.int64)
    xgoal = np.asarray(xgoal, dtype=np.int64)
    pts, cst, par = _fresh(n, xstart)
    seen = set()
    cnt = _Counter()
    acc = np.zeros(n, dtype=bool) if trace else None
    j = 1
    for i in range(n):
        xnew = samples[i]
        vnear = nearest(pts, xnew)
        ok = cnt.free(og, pts[vnear], xnew)
        key = (int(xnew[0]), int(xnew[1]))
        if ok and key not in seen and j != n:                 # rrt.py:425
            seen.add(key)
            pts[j] = xnew
            cst[j] = edge_cost(cst, pts, vnear, xnew)
            par[j] = vnear
            if trace:
                acc[i] = True
            j += 1
    vgoal, found, pts, cst, par = connect_goal(og, cst, pts, par, xgoal, j, cnt)
    return Tree(pts, cst, par, j, vgoal, found, n, cnt.checks, cnt.cells, accepted=acc)


def _choose_parent(og, pts, cst, xnew, vnear, ring, cnt):
    """rrt.py:510-521 -- running strict-< minimum over the radius set in ascending index
    order, collision test only for candidates that would improve."""
    vbest = vnear
    cbest = edge_cost(cst, pts, vnear, xnew)
    for vn in ring:
        cn = edge_cost(cst, pts, vn, xnew)
        if cn < cbest:
            if cnt.free(og, pts[vn], xnew):
                vbest = int(vn)
                cbest = cn
    return vbest, cbest


def _dead_rewire(pts, cst, xnew, ring) -> int:
    """rrt.py:532-546 -- count how often ``cost(vn -> xnew) < vcosts[vn]`` holds (never)."""
    fired = 0
    for vn in ring:
        if edge_cost(cst, pts, vn, xnew) < cst[vn]:
            fired += 1
    return fired


def plan_star(og, n, r_rewire, xstart, xgoal, samples, trace=False, count_rewire=False) -> Tree:
    """RRTStar.plan on an explicit sample stream (rrt.py:466-556)."""
    xstart = np.asarray(xstart, dtype=np.int64)
    xgoal = np.asarray(xgoal, dtype=np.int64)
    pts, cst, par = _fresh(n, xstart)
    seen = set()
    cnt = _Counter()
    acc = np.zeros(n, dtype=bool) if trace else None
    fired = 0
    j = 1
    for i in range(n):
        xnew = samples[i]
        vnear = nearest(pts, xnew)
        ok = cnt.free(og, pts[vnear], xnew)
        key = (int(xnew[0]), int(xnew[1]))
        if ok and key not in seen and j != n:                 # rrt.py:507
            seen.add(key)
            ring = within(pts, xnew, r_rewire)
            ring = ring[ring < j]      # unfilled rows can alias in (int64 wrap); their cost is inf
            vbest, cbest = _choose_parent(og, pts, cst, xnew, vnear, ring, cnt)
            pts[j] = xnew
            cst[j] = cbest
            par[j] = vbest
            if count_rewire:
                fired += _dead_rewire(pts, cst, xnew, ring)
            if trace:
                acc[i] = True
            j += 1
    vgoal, found, pts, cst, par = connect_goal(og, cst, pts, par, xgoal, j, cnt)
    return Tree(pts, cst, par, j, vgoal, found, n, cnt.checks, cnt.cells, fired, accepted=acc)


def ellipse_rotation(xstart, xgoal) -> np.ndarray:
    """2x2 world-frame rotation of the informed ellipse, computed with the same numpy calls as
    rrt.py:601-613 (SVD of the outer product; sign convention is LAPACK's)."""
    xstart = np.asarray(xstart)
    xgoal = np.asarray(xgoal)
    axis = np.atleast_2d((xgoal - xstart) / np.linalg.norm(xgoal - xstart))
    m = np.outer(axis, np.atleast_2d([1, 0]))
    u, _, v = np.linalg.svd(m)
    return u @ np.diag([np.linalg.det(u), np.linalg.det(v)]) @ v.T


def unitball_from_uniform(u1: float, u2: float) -> np.ndarray:
    """rrt.py:579-587 with the two uniform draws supplied (radius draw first)."""
    theta = 2 * np.pi * u2
    return np.array([np.sqrt(u1) * np.cos(theta), np.sqrt(u1) * np.sin(theta)])


def ellipse_sample(shape, rot, xstart, xgoal, c, ball) -> np.ndarray:
    """rrt.py:589-599 + 615-625: scale the unit-ball point by (c/2, sqrt(|c^2-d^2|)/2), rotate,
    centre, clamp to the grid and truncate to int."""
    centre = (xstart + xgoal) / 2
    r1 = c / 2
    gap = xstart - xgoal
    d2 = np.dot(gap.T, gap)
    r2 = np.sqrt(abs(c * c - d2)) / 2
    cl = np.dot(rot, np.diag([r1, r2]))
    x, y = tuple(np.dot(cl, ball) + centre)
    x = int(max(0, min(shape[0] - 1, x)))
    y = int(max(0, min(shape[1] - 1, y)))
    return np.array((x, y))


def plan_informed(og, n, r_rewire, r_goal, xstart, xgoal, samples, balls,
                  rot: Optional[np.ndarray] = None, trace=False) -> Tree:
    """RRTStarInformed.plan (rrt.py:653-758) on explicit streams: ``samples[i]`` is used while no
    solution vertex exists, ``balls[i]`` (unit-ball points, rrt.py:579-587) afterwards.  Stream
    entries are indexed by iteration, so exactly one of the two is consumed per iteration."""
    xstart = np.asarray(xstart, dtype=np.int64)
    xgoal = np.asarray(xgoal, dtype=np.int64)
    pts, cst, par = _fresh(n, xstart)
    seen = set()
    cnt = _Counter()
    acc = np.zeros(n, dtype=bool) if trace else None
    soln = []
    ell: Dict[int, float] = {}
    first_iter, ell_iters = -1, 0
    j = 1
    for i in range(n):
        if not soln:
            xnew = samples[i]
        else:
            if rot is None:
                rot = ellipse_rotation(xstart, xgoal)
            k = int(np.argmin(cst[soln])) if len(soln) > 1 else 0          # rrt.py:627-633
            c = cst[soln[k]] + r2norm(xgoal - pts[soln[k]])                # rrt.py:698-699
            xnew = ellipse_sample(og.shape, rot, xstart, xgoal, c, balls[i])
            ell[j] = float(c)
            ell_iters += 1
        vnear = nearest(pts, xnew)
        ok = cnt.free(og, pts[vnear], xnew)
        key = (int(xnew[0]), int(xnew[1]))
        if ok and key not in seen and j != n:                 # rrt.py:707
            seen.add(key)
            ring = within(pts, xnew, r_rewire)
            ring = ring[ring < j]
            vbest, cbest = _choose_parent(og, pts, cst, xnew, vnear, ring, cnt)
            pts[j] = xnew
            cst[j] = cbest
            par[j] = vbest
            if r2norm(xnew - xgoal) < r_goal:                 # rrt.py:744-745
                if not soln:
                    first_iter = i
                soln.append(j)
            if trace:
                acc[i] = True
            j += 1
    vgoal, found, pts, cst, par = connect_goal(og, cst, pts, par, xgoal, j, cnt)
    return Tree(pts, cst, par, j, vgoal, found, n, cnt.checks, cnt.cells, 0,
                first_iter, ell_iters, ell, acc)


# --------------------------------------------------------------------------------------
# graph shape of the public API (rrt.py:334-369, 87-129) -- used to check the drop-in classes
# --------------------------------------------------------------------------------------
def graph_records(tree: Tree):
    """(node_order, edges) as a networkx DiGraph built by rrt.py:359-368 enumerates them: the goal
    node first, then one node per row; edges (parent, child, dist, cost) grouped by parent in node
    order, children of one parent in insertion (= child index) order."""
    rows = tree.points.shape[0]
    nodes = [int(tree.vgoal)] + [i for i in range(rows) if i != int(tree.vgoal)]
    kids = {}
    for child in range(rows):
        p = int(tree.parents[child])
        if p >= 0:
            d = r2norm(tree.points[child] - tree.points[p])
            kids.setdefault(p, []).append((p, child, d, float(tree.vcosts[child])))
    edges = [e for u in nodes for e in kids.get(u, [])]
    return nodes, edges
