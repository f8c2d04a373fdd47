"""Replanning on a time-varying grid: the frame loop of the reference's
``DynamicEnvironmentAnimation.simulate_dynamic_goals`` (anim.py:56-115) without its matplotlib half.

Per frame the reference (anim.py:69-113):

1. draws a new goal when the agent is within ``r_within_goal`` of the current one (anim.py:72-73),
2. clamps the position to the grid (anim.py:74),
3. inflates the obstacles by ``bsize = int(movespeed / 2)`` 4-connected dilation steps, keeps the
   buffer away from the square ``[p, p + 2 bsize)^2`` the agent stands in and ORs it into the grid
   (anim.py:79-87)  ->  ``rrtk_inflate_grid``,
4. ``rrtobj.set_og(og)``; plans from the position to the goal with the SAME planner object, whose
   random generator keeps running (rrt.py:85)  ->  ``rrtk_free_rows`` + ``rrtk_sample_streams_carry``
   + ``rrtk_plan_batch``; the shipped code calls the removed ``make`` / ``path_points``
   (anim.py:93-94), meant are ``plan`` and ``vertices_as_ndarray``,
5. moves ``movespeed`` cells along the first path segment unless it stands inside an obstacle or no
   path exists (anim.py:107-113).

``BatchReplanner`` runs that loop for many agents at once (one plan per agent and frame, all agents
of a frame in one kernel launch, grids never leave the GPU); only the first path segment of every
agent comes back to the host each frame.  ``DynamicEnvironment`` is the single-agent form on top
of the drop-in planner classes, with the reference's constructor.  Results are pinned by
``oracle/replan_oracle.py`` (scipy's binary_dilation + the plan oracle) in tests/test_gpu_replan.py.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _lib, batch
from .rrt import RRT, random_point_og


def clamp(xy, shape) -> np.ndarray:
    """anim.py:50-54."""
    return np.array([max(0, min(int(xy[0]), shape[0] - 1)), max(0, min(int(xy[1]), shape[1] - 1))])


def step_along(position: np.ndarray, first_segment: np.ndarray, movespeed: float) -> np.ndarray:
    """anim.py:107-113: move ``movespeed`` along the first path segment, truncating to integers."""
    v = first_segment[1] - first_segment[0]
    angle = np.arctan2(v[1], v[0])
    return position + (np.array([np.cos(angle), np.sin(angle)]) * movespeed).astype(np.int64)


class DynamicEnvironment(object):
    """Single agent, reference constructor (anim.py:25-48) minus ``appearance``'s use."""

    def __init__(self, movespeed: float, r_within_goal: float, rrtobj: RRT, appearance: dict = None, buffer_size: int = None):
        self.movespeed = movespeed
        self.r_within_goal = r_within_goal
        self.rrtobj = rrtobj
        self.buffer_size = buffer_size if buffer_size is not None else int(movespeed / 2)
        self.appearance = appearance

    clamp = staticmethod(clamp)

    def inflate(self, og: np.ndarray, position: np.ndarray) -> np.ndarray:
        """anim.py:79-87 on the device (K0 + rrtk_inflate_grid + unpack)."""
        bsize = int(self.movespeed / 2)          # the reference ignores buffer_size here (anim.py:79)
        if bsize < 1:
            raise ValueError("movespeed < 2 makes the reference dilate until nothing changes (scipy iterations=0); not supported")
        ctx = self.rrtobj._device()
        return ctx.inflate((np.asarray(og) != 0).astype(np.uint8), bsize, np.array([[position[0], position[1], 2 * bsize]]))[0].astype(np.asarray(og).dtype)

    def simulate_dynamic_goals(self, og_3d: np.ndarray, rnd_gen: Optional[np.random.Generator] = None):
        """Returns (goals, positions, paths, trees) like the reference (anim.py:56-115)."""
        frames = og_3d.shape[0]
        xstart = random_point_og(og_3d[0], rnd_gen)
        xgoal = random_point_og(og_3d[0], rnd_gen)
        current_position = np.array(xstart, dtype=np.int64)
        goals, positions = np.empty((frames, 2)), np.empty((frames, 2))
        paths, trees = [], []
        for fi in range(frames):
            og = og_3d[fi]
            if np.linalg.norm(current_position - xgoal) < self.r_within_goal:
                xgoal = random_point_og(og, rnd_gen)
            current_position = clamp(current_position, og.shape)
            og = self.inflate(og, current_position)
            in_obs = og[current_position[0], current_position[1]] == 1
            self.rrtobj.set_og(og)
            T, gv = self.rrtobj.plan(current_position, xgoal)
            path = self.rrtobj.vertices_as_ndarray(T, self.rrtobj.route2gv(T, gv))
            goals[fi] = xgoal
            positions[fi] = current_position
            paths.append(path)
            trees.append(np.array([[T.nodes[a]["pt"], T.nodes[b]["pt"]] for a, b in T.edges]))
            if path.shape[0] != 0 and not in_obs:
                current_position = step_along(current_position, path[0], self.movespeed)
        return goals, positions, paths, trees


class BatchReplanner:
    """Many agents on one sequence of frames, device resident (see the module docstring).

    ``kind`` in {"standard", "star"}; agent a plans with ``default_rng(seeds[a])`` as its planner object's
    generator and ``goal_rngs[a]`` for the goals it draws (anim.py:60-61, 73)."""

    PATH_CAP = 1024

    def __init__(self, kind: str, W: int, H: int, n: int, movespeed: float, r_within_goal: float, r_rewire: float = 0.0,
                 device: Optional[int] = None, threads: int = 0):
        if kind not in ("standard", "star"):
            raise ValueError("BatchReplanner plans with RRTStandard or RRTStar")
        self.bsize = int(movespeed / 2)
        if self.bsize < 1:
            raise ValueError("movespeed must be at least 2 (anim.py:79: bsize = int(movespeed / 2) dilation steps)")
        self.movespeed, self.r_within_goal = movespeed, r_within_goal
        self.db = batch.DeviceBatch(kind, W, H, n, r_rewire, device=device, threads=threads)
        self.W, self.H, self.n = W, H, n

    def simulate(self, frames_u8, starts: np.ndarray, goals: np.ndarray, seeds: Sequence[int], goal_rngs: Sequence[np.random.Generator],
                 keep_trees: bool = False):
        """frames_u8: (F, W, H) uint8 grids (numpy or CUDA tensor).  starts / goals: (A, 2) ints.
        Returns dict(positions (F, A, 2), goals (F, A, 2), found (F, A), path_cost (F, A), first_segment (F, A, 2, 2),
        [trees: list of BatchResult])."""
        db, t = self.db, self.db.torch
        L = db.L
        A = int(np.asarray(starts).shape[0])
        fr = frames_u8 if t.is_tensor(frames_u8) else t.from_numpy(np.ascontiguousarray(frames_u8, dtype=np.uint8))
        fr = fr.to(db.dev)
        F = int(fr.shape[0])
        words = db.words
        frame_bits = db._empty((F, words), t.int32)
        frame_rowcum = db._empty((F, self.W + 1), t.int32)
        with t.cuda.device(db.dev):
            _lib.check(L.rrtk_pack_grid(fr.data_ptr(), F, self.W, self.H, frame_bits.data_ptr(), db._stream()), "pack")
            _lib.check(L.rrtk_free_rows(frame_bits.data_ptr(), F, self.W, self.H, frame_rowcum.data_ptr(), db._stream()), "free_rows")
        frame_free = None                                       # host free lists, made lazily when a goal must be redrawn
        db.bits = db._empty((A, words), t.int32)
        db.rowcum = db._empty((A, self.W + 1), t.int32)
        scratch = db._empty((2, words), t.int32)
        state = t.from_numpy(batch.seed_states(seeds).view(np.int64)).to(db.dev)
        carry = t.zeros((A, 2), dtype=t.int32, device=db.dev)
        db.samples = db._empty((A, self.n, 2), t.int16)
        pos = np.array(starts, dtype=np.int64).reshape(A, 2)
        goal = np.array(goals, dtype=np.int64).reshape(A, 2)
        out = dict(positions=np.empty((F, A, 2), dtype=np.int64), goals=np.empty((F, A, 2), dtype=np.int64),
                   found=np.zeros((F, A), dtype=bool), path_cost=np.full((F, A), np.inf), first_segment=np.zeros((F, A, 2, 2), dtype=np.int64),
                   in_obstacle=np.zeros((F, A), dtype=bool))
        trees = []
        for fi in range(F):
            reached = np.linalg.norm((pos - goal).astype(np.float64), axis=1) < self.r_within_goal
            if reached.any():
                if frame_free is None:
                    frame_free = {}
                if fi not in frame_free:
                    frame_free[fi] = np.argwhere(fr[fi].cpu().numpy() == 0)
                for a in np.flatnonzero(reached):
                    free = frame_free[fi]
                    goal[a] = free[goal_rngs[a].integers(low=0, high=free.shape[0])]
            pos = np.stack([clamp(p, (self.W, self.H)) for p in pos])
            holes = t.from_numpy(np.concatenate([pos, np.full((A, 1), 2 * self.bsize)], axis=1).astype(np.int32)).to(db.dev)
            with t.cuda.device(db.dev):
                _lib.check(L.rrtk_inflate_grid(frame_bits[fi].data_ptr(), 1, self.W, self.H, self.bsize, holes.data_ptr(), A,
                                               db.bits.data_ptr(), scratch.data_ptr(), db._stream()), "inflate_grid")
                _lib.check(L.rrtk_free_rows(db.bits.data_ptr(), A, self.W, self.H, db.rowcum.data_ptr(), db._stream()), "free_rows")
            db.set_plans(batch.make_desc(np.arange(A), pos, goal))
            with t.cuda.device(db.dev):
                _lib.check(L.rrtk_sample_streams_carry(db.bits.data_ptr(), db.rowcum.data_ptr(), self.W, self.H, db.desc.data_ptr(), A,
                                                       state.data_ptr(), carry.data_ptr(), self.n, db.samples.data_ptr(), db._stream()),
                           "sample_streams_carry")
            db.run()
            path, ln = db.paths(self.PATH_CAP)                  # root -> goal vertex ids (rrtk_extract_paths)
            # the agent's own cell after inflation (anim.py:88-91)
            segs = t.from_numpy(np.concatenate([pos, pos], axis=1).astype(np.int32)).to(db.dev)
            wid = t.arange(A, dtype=t.int32, device=db.dev)
            free_cell = db._empty((A,), t.uint8)
            with t.cuda.device(db.dev):
                _lib.check(L.rrtk_collision_segments(db.bits.data_ptr(), self.W, self.H, segs.data_ptr(), wid.data_ptr(), A,
                                                     free_cell.data_ptr(), None, db._stream()), "collision_segments")
            stats = db.out["stats"].cpu().numpy()
            vgoal = stats[:, 1]
            found = stats[:, 2] != 0
            ln_h = ln.cpu().numpy()
            if (ln_h > self.PATH_CAP).any():
                raise RuntimeError("a path is deeper than BatchReplanner.PATH_CAP vertices")
            # only the second vertex of every path and the goal's cost come back to the host
            rows = t.arange(A, device=db.dev)
            second = t.where(ln >= 2, path[:, 1], t.zeros_like(ln)).long()
            pts_h = db.out["pts"][rows, second].cpu().numpy().astype(np.int64)
            cost_h = db.out["cost"][rows, t.from_numpy(np.where(found, vgoal, 0)).to(db.dev)].cpu().numpy()
            found = found & (ln_h >= 2)
            in_obs = free_cell.cpu().numpy() == 0
            out["positions"][fi], out["goals"][fi], out["found"][fi], out["in_obstacle"][fi] = pos, goal, found, in_obs
            out["path_cost"][fi] = np.where(found, cost_h, np.inf)
            out["first_segment"][fi, :, 0], out["first_segment"][fi, :, 1] = pos, pts_h
            if keep_trees:
                trees.append(db.download())
            for a in range(A):
                if found[a] and not in_obs[a]:
                    pos[a] = step_along(pos[a], np.stack([pos[a], pts_h[a]]), self.movespeed)
        if keep_trees:
            out["trees"] = trees
        return out

