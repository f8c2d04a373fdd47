"""ctypes front-end of the C oracle (oracle/rrt_oracle.c) -- TEST INFRASTRUCTURE ONLY.

``build()`` compiles ``oracle/rrt_oracle.c`` with gcc into ``oracle/_build/liboracle.so`` (git-ignored,
travels to the GPU box with the snapshot).  Nothing in ``rrtplanner_b200`` imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .rrt_oracle import Tree, UNFILLED

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "rrt_oracle.c")
_OUT = os.path.join(_HERE, "_build", "liboracle.so")

ST_NAMES = ("j", "vgoal", "found", "checks", "cells", "first_solution_iter", "ellipse_iters",
            "nn_pairs", "ring_members", "accepted", "rewire_fired")
KINDS = {"standard": 0, "star": 1, "informed": 2}

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_OUT):
            build()
        L = C.CDLL(_OUT)
        L.orc_first_hit.restype = C.c_int
        L.orc_first_hit.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.orc_collision_batch.restype = None
        L.orc_collision_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        L.orc_nearest.restype = C.c_int
        L.orc_nearest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_within.restype = C.c_int
        L.orc_within.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.orc_plan.restype = C.c_int
        L.orc_plan.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 10
        _lib = L
    return _lib


def _u8(og):
    return np.ascontiguousarray(og != 0, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def collision_batch(og, segs):
    """(free[nseg] bool, cells[nseg] int32) for int segments (ax, ay, bx, by) -- rrt.py:183-229."""
    g = _u8(og)
    s = np.ascontiguousarray(segs, dtype=np.int32)
    free = np.empty(s.shape[0], dtype=np.uint8)
    cells = np.empty(s.shape[0], dtype=np.int32)
    lib().orc_collision_batch(_p(g), g.shape[0], g.shape[1], _p(s), s.shape[0], _p(free), _p(cells))
    return free.astype(bool), cells


def nearest(points, j, x):
    p = np.ascontiguousarray(points[:j], dtype=np.int32)
    d2 = C.c_int64(0)
    v = lib().orc_nearest(_p(p), j, int(x[0]), int(x[1]), C.byref(d2))
    return int(v), int(d2.value)


def within(points, j, x, r):
    p = np.ascontiguousarray(points[:j], dtype=np.int32)
    out = np.empty(max(j, 1), dtype=np.int32)
    m = lib().orc_within(_p(p), j, int(x[0]), int(x[1]), float(r), _p(out))
    return out[:m].astype(np.int64)


def plan_raw(kind, og, n, xstart, xgoal, samples, r_rewire=0.0, r_goal=0.0, balls=None, rot=None):
    """Device-layout result: pts (n+1,2) int32 (INT32_MIN = unfilled), cost (n+1,), parent (n+1,),
    stats dict, ell_c (n+1,)."""
    g = _u8(og)
    s = np.ascontiguousarray(samples, dtype=np.int32)
    assert s.shape == (n, 2)
    st = np.ascontiguousarray(xstart, dtype=np.int32)
    gl = np.ascontiguousarray(xgoal, dtype=np.int32)
    b = None if balls is None else np.ascontiguousarray(balls, dtype=np.float64)
    rt = None if rot is None else np.ascontiguousarray(rot, dtype=np.float64).reshape(4)
    pts = np.empty((n + 1, 2), dtype=np.int32)
    cost = np.empty(n + 1)
    par = np.empty(n + 1, dtype=np.int32)
    stats = np.zeros(len(ST_NAMES), dtype=np.int64)
    ell = np.empty(n + 1)
    rc = lib().orc_plan(KINDS[kind], _p(g), g.shape[0], g.shape[1], n, float(r_rewire), float(r_goal),
                        _p(st), _p(gl), _p(s), _p(b), _p(rt), _p(pts), _p(cost), _p(par), _p(stats), _p(ell))
    if rc:
        raise MemoryError("orc_plan")
    return pts, cost, par, dict(zip(ST_NAMES, (int(v) for v in stats))), ell


def rows_like_reference(pts, cost, par, j, vgoal, found, n):
    """Expand the compact (n+1)-row device/C layout into the reference's arrays after go2goal
    (rrt.py:320-323): n rows, or n + 1 rows with row n duplicating the goal when it was found."""
    rows = n + 1 if found else n
    P = np.full((rows, 2), UNFILLED, dtype=np.int64)
    Cc = np.full((rows,), np.inf)
    Pa = np.full((rows,), -1, dtype=np.int64)
    top = j + 1 if found else j
    P[:top] = pts[:top]
    Cc[:top] = cost[:top]
    Pa[:top] = par[:top]
    if found and j < n:
        P[n] = pts[j]
        Cc[n] = cost[j]
    return P, Cc, Pa


def plan(kind, og, n, xstart, xgoal, samples, r_rewire=0.0, r_goal=0.0, balls=None, rot=None) -> Tree:
    pts, cost, par, st, ell = plan_raw(kind, og, n, xstart, xgoal, samples, r_rewire, r_goal, balls, rot)
    P, Cc, Pa = rows_like_reference(pts, cost, par, st["j"], st["vgoal"], bool(st["found"]), n)
    ell_c = {int(k): float(ell[k]) for k in np.flatnonzero(~np.isnan(ell))}
    return Tree(P, Cc, Pa, st["j"], st["vgoal"], bool(st["found"]), n, st["checks"], st["cells"],
                st["rewire_fired"], st["first_solution_iter"], st["ellipse_iters"], ell_c)
