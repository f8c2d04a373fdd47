"""CPU restatement of the reference's replanning loop -- TEST INFRASTRUCTURE ONLY.

Follows ``DynamicEnvironmentAnimation.simulate_dynamic_goals`` (/root/reference/rrtplanner/anim.py:56-115)
line by line, with

* scipy.ndimage.binary_dilation itself for the obstacle buffer (anim.py:80; scipy is the reference's own
  dependency, setup.py:123),
* the plan oracle of oracle/rrt_oracle.c for ``rrtobj.make`` (the shipped call, anim.py:93, names a
  method that no longer exists; ``plan`` is meant -- SURVEY.md section 0, quirk 7) with the planner
  object's generator running on from frame to frame (rrt.py:85, 231-240),
* explicit generators where the reference uses the global one (``random_point_og`` without ``rnd_gen``).

Nothing in rrtplanner_b200 imports this module.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import binary_dilation

from . import c_oracle


def clamp(xy, shape):
    return np.array([max(0, min(int(xy[0]), shape[0] - 1)), max(0, min(int(xy[1]), shape[1] - 1))])


def inflate(og: np.ndarray, position, bsize: int) -> np.ndarray:
    """anim.py:79-87."""
    og = (np.asarray(og) != 0).astype(np.int64)
    og_dilated = binary_dilation(og, iterations=bsize)
    buffer_reg = og_dilated - og
    for i in range(bsize * 2):
        for j in range(bsize * 2):
            hole = clamp(np.asarray(position) + np.array([i, j]), og.shape)
            buffer_reg[hole[0], hole[1]] = 0
    return og | buffer_reg


def simulate(kind: str, og_3d: np.ndarray, n: int, r_rewire: float, movespeed: float, r_within_goal: float,
             xstart, xgoal, seed: int, goal_rng: np.random.Generator):
    """One agent.  Returns per-frame records: position, goal, inflated grid, tree (pts, cost, parent, stats), moved."""
    rand_gen = np.random.default_rng(seed)                    # the planner object's generator (rrt.py:85)
    current_position = np.array(xstart, dtype=np.int64)
    xgoal = np.array(xgoal, dtype=np.int64)
    recs = []
    for fi in range(og_3d.shape[0]):
        og = og_3d[fi]
        if np.linalg.norm(current_position - xgoal) < r_within_goal:
            free0 = np.argwhere(og == 0)
            xgoal = free0[goal_rng.integers(low=0, high=free0.shape[0])]
        current_position = clamp(current_position, og.shape)
        bsize = int(movespeed / 2)
        og = inflate(og, current_position, bsize)
        in_obs = og[current_position[0], current_position[1]] == 1
        free = np.argwhere(og == 0)                           # set_og (rrt.py:261-272)
        samples = free[rand_gen.integers(0, free.shape[0], size=n)]
        pts, cost, par, st, _ = c_oracle.plan_raw(kind, og, n, current_position, xgoal, samples, r_rewire)
        rec = dict(position=current_position.copy(), goal=xgoal.copy(), og=og.astype(np.uint8), pts=pts, cost=cost, parent=par, stats=st,
                   in_obs=bool(in_obs), samples=samples)
        # path root -> goal (route2gv on a tree = parent walk); first segment drives the motion (anim.py:107-113)
        moved = False
        if st["found"]:
            v = st["vgoal"]
            while par[v] > 0:
                v = par[v]
            if par[v] == 0 and not in_obs:
                seg = np.stack([current_position, pts[v].astype(np.int64)])
                d = seg[1] - seg[0]
                angle = np.arctan2(d[1], d[0])
                current_position = current_position + (np.array([np.cos(angle), np.sin(angle)]) * movespeed).astype(np.int64)
                moved = True
        rec["moved"] = moved
        recs.append(rec)
    return recs
