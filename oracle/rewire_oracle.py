"""ctypes front-end of oracle/rewire_oracle.c -- TEST INFRASTRUCTURE ONLY.

The C file is the specification (parity status: UNPINNED, see its header) of the planners the reference
advertises but does not contain: RRT* with a rewire step that fires, and the Dubins-vehicle RRT / RRT*
on the Dubins primitive (README.md:12,18-19).  Nothing in ``rrtplanner_b200`` imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "rewire_oracle.c")
_OUT = os.path.join(_HERE, "_build", "liboracle2.so")

MODELS = {"euclid": 0, "dubins": 1}
ST_NAMES = ("j", "vgoal", "found", "checks", "accepted", "rewires", "propagated", "ring_members", "len_evals", "overflow",
            "ell_iters", "first_solution_iter")
WORDS = ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_OUT)
        vp, i, d, lg = C.c_void_p, C.c_int, C.c_double, C.c_long
        L.orc2_dubins_batch.restype = None
        L.orc2_dubins_batch.argtypes = [vp, lg, i, d, vp, vp, vp]
        L.orc2_dubins_free_batch.restype = None
        L.orc2_dubins_free_batch.argtypes = [vp, i, i, vp, lg, i, d, d, vp]
        L.orc2_dubins_points.restype = None
        L.orc2_dubins_points.argtypes = [vp, i, d, vp, lg, vp]
        L.orc2_dubins_all.restype = None
        L.orc2_dubins_all.argtypes = [vp, i, d, vp, vp]
        L.orc2_math.restype = None
        L.orc2_math.argtypes = [vp, vp, lg, vp, vp, vp]
        L.orc2_plan.restype = i
        L.orc2_plan.argtypes = [i, i, i, vp, i, i, i, d, i, d, d] + [vp] * 9
        L.orc2_plan_informed.restype = i
        L.orc2_plan_informed.argtypes = [i, i, i, vp, i, i, i, d, i, d, d] + [vp] * 9 + [d, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(og):
    return np.ascontiguousarray(np.asarray(og) != 0, dtype=np.uint8)


def math(a, b):
    """(dm_atan2(a, b), dm_sin(a), dm_cos(a)) elementwise."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    at2, sn, cs = np.empty_like(a), np.empty_like(a), np.empty_like(a)
    lib().orc2_math(_p(a), _p(b), a.size, _p(at2), _p(sn), _p(cs))
    return at2, sn, cs


def dubins(q, nh, rho):
    """q: (nq, 6) int (x0, y0, h0, x1, y1, h1) -> (word int32, tpq (nq,3), length)."""
    q = np.ascontiguousarray(q, dtype=np.int32).reshape(-1, 6)
    word = np.empty(q.shape[0], dtype=np.int32)
    tpq = np.empty((q.shape[0], 3))
    ln = np.empty(q.shape[0])
    lib().orc2_dubins_batch(_p(q), q.shape[0], int(nh), float(rho), _p(word), _p(tpq), _p(ln))
    return word, tpq, ln


def dubins_all(q, nh, rho):
    q = np.ascontiguousarray(q, dtype=np.int32).reshape(6)
    ok = np.empty(6, dtype=np.int32)
    tpq = np.zeros((6, 3))
    lib().orc2_dubins_all(_p(q), int(nh), float(rho), _p(ok), _p(tpq))
    return ok.astype(bool), tpq


def dubins_free(og, q, nh, rho, ds):
    g = _u8(og)
    q = np.ascontiguousarray(q, dtype=np.int32).reshape(-1, 6)
    out = np.empty(q.shape[0], dtype=np.uint8)
    lib().orc2_dubins_free_batch(_p(g), g.shape[0], g.shape[1], _p(q), q.shape[0], int(nh), float(rho), float(ds), _p(out))
    return out.astype(bool)


def dubins_points(q, nh, rho, s):
    """(x, y, theta) of the shortest path of ONE query at arc lengths s."""
    q = np.ascontiguousarray(q, dtype=np.int32).reshape(6)
    s = np.ascontiguousarray(s, dtype=np.float64)
    out = np.empty((s.size, 3))
    lib().orc2_dubins_points(_p(q), int(nh), float(rho), _p(s), s.size, _p(out))
    return out


def plan(model, og, n, start, goal, samples, star=True, rewire=True, r_rewire=0.0, nh=16, rho=1.0, ds=1.0, informed=None):
    """start / goal: (x, y, h); samples: (n, 3) (x, y, h) (h ignored by the Euclidean model).
    ``informed``: None, or dict(r_goal, rot (2, 2), balls (n, 2) or None for the probe) -- the informed sampling rule.
    Returns dict(pts (n+1,2) int32, head, cost, elen, parent, stats dict[, ell (n+1)])."""
    g = _u8(og)
    s = np.ascontiguousarray(samples, dtype=np.int32)
    assert s.shape == (n, 3), s.shape
    st = np.ascontiguousarray(start, dtype=np.int32).reshape(3)
    gl = np.ascontiguousarray(goal, dtype=np.int32).reshape(3)
    pts = np.empty((n + 1, 2), dtype=np.int32)
    head = np.empty(n + 1, dtype=np.int32)
    cost, elen = np.empty(n + 1), np.empty(n + 1)
    par = np.empty(n + 1, dtype=np.int32)
    stats = np.zeros(len(ST_NAMES), dtype=np.int64)
    args = (MODELS[model], int(bool(star)), int(bool(rewire)), _p(g), g.shape[0], g.shape[1], n, float(r_rewire),
            int(nh), float(rho), float(ds), _p(st), _p(gl), _p(s), _p(pts), _p(head), _p(cost), _p(elen), _p(par), _p(stats))
    ell = None
    if informed is None:
        rc = lib().orc2_plan(*args)
    else:
        rot = np.ascontiguousarray(informed["rot"], dtype=np.float64).reshape(4)
        balls = None if informed.get("balls") is None else np.ascontiguousarray(informed["balls"], dtype=np.float64).reshape(n, 2)
        ell = np.empty(n + 1)
        rc = lib().orc2_plan_informed(*args, float(informed["r_goal"]), _p(rot), _p(balls), _p(ell))
    if rc:
        raise MemoryError("orc2_plan")
    out = dict(pts=pts, head=head, cost=cost, elen=elen, parent=par, stats=dict(zip(ST_NAMES, (int(v) for v in stats))))
    if ell is not None:
        out["ell"] = ell
    return out
