"""Replanning caller (anim.py:56-115): device inflation == scipy's
binary_dilation recipe, generator carry == numpy, and the batched frame loop == the CPU restatement
agent by agent (positions, goals, inflated grids, trees bit for bit)."""
import numpy as np
import pytest

from oracle import replan_oracle as RO
from rrtplanner_b200 import _lib, batch, replan, rrt, worlds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(64, 64), (43, 100), (100, 43), (33, 31), (130, 70)])
@pytest.mark.parametrize("iters", [1, 2, 5])
def test_inflate_equals_scipy_recipe(shape, iters):
    rng = np.random.default_rng(shape[0] * 7 + iters)
    og = (rng.random(shape) < 0.08).astype(np.uint8)
    og[0, :3] = 1; og[-1, -1] = 1                       # obstacles on the border: nothing may leak in from outside
    holes = np.array([[0, 0, 2 * iters], [shape[0] - 2, shape[1] - 3, 2 * iters], [shape[0] // 2, shape[1] // 2, 2 * iters], [5, 7, 0]])
    ctx = _lib.Context()
    got = ctx.inflate(og, iters, holes)
    for k, (px, py, sz) in enumerate(holes):
        want = RO.inflate(og, (px, py), iters) if sz else (og | RO.binary_dilation(og, iterations=iters)).astype(np.int64)
        assert np.array_equal(got[k], want.astype(np.uint8)), (shape, iters, k)
    assert np.array_equal(ctx.inflate(og, iters)[0], (og | RO.binary_dilation(og, iterations=iters)).astype(np.uint8))
    ctx.close()


def test_generator_carry_matches_numpy_across_calls():
    import torch
    W = H = 96
    og = worlds.perlin_occupancygrid(W, H, seed=5).astype(np.uint8)
    A, n = 6, 333                                           # odd count: a buffered 32-bit half is carried over
    db = batch.DeviceBatch("standard", W, H, n).set_worlds_host(og[None])
    db.set_plans(batch.make_desc(np.zeros(A, int), np.zeros((A, 2)), np.zeros((A, 2))))
    seeds = np.arange(40, 40 + A)
    state = torch.from_numpy(batch.seed_states(seeds).view(np.int64)).cuda()
    carry = torch.zeros((A, 2), dtype=torch.int32, device="cuda")
    smp = torch.empty((A, n, 2), dtype=torch.int16, device="cuda")
    free = np.argwhere(og == 0)
    gens = [np.random.default_rng(int(s)) for s in seeds]
    for call in range(4):
        _lib.check(db.L.rrtk_sample_streams_carry(db.bits.data_ptr(), db.rowcum.data_ptr(), W, H, db.desc.data_ptr(), A, state.data_ptr(),
                                                  carry.data_ptr(), n, smp.data_ptr(), torch.cuda.current_stream().cuda_stream), "carry")
        got = smp.cpu().numpy().astype(np.int64)
        for a in range(A):
            assert np.array_equal(got[a], free[gens[a].integers(0, free.shape[0], size=n)]), (call, a)


@pytest.mark.parametrize("kind", ["standard", "star"])
def test_batched_frame_loop_equals_restatement(kind):
    W = H = 128
    F, A, n, speed, rgoal = 6, 5, 300, 6, 14.0
    og_3d = worlds.perlin_occupancygrid(W, H, frames=F, seed=21).astype(np.uint8)
    assert og_3d.shape == (F, W, H)
    pairs = [worlds.start_goal(og_3d[0], 50 + a) for a in range(A)]
    starts, goals = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    goals[1] = starts[1] + np.array([3, 2])              # agent 1 starts inside the goal radius: draws a new goal at once
    seeds = 900 + np.arange(A)
    rp = replan.BatchReplanner(kind, W, H, n, speed, rgoal, r_rewire=20.0)
    out = rp.simulate(og_3d, starts, goals, seeds, [np.random.default_rng(7000 + a) for a in range(A)], keep_trees=True)
    moved_any = False
    for a in range(A):
        recs = RO.simulate(kind, og_3d, n, 20.0, speed, rgoal, starts[a], goals[a], int(seeds[a]), np.random.default_rng(7000 + a))
        for fi, rec in enumerate(recs):
            assert np.array_equal(out["positions"][fi, a], rec["position"]), (a, fi)
            assert np.array_equal(out["goals"][fi, a], rec["goal"]), (a, fi)
            assert bool(out["in_obstacle"][fi, a]) == rec["in_obs"]
            tr, st = out["trees"][fi], rec["stats"]
            top = st["j"] + (1 if st["found"] else 0)
            assert int(tr.stats[a, 0]) == st["j"] and int(tr.stats[a, 1]) == st["vgoal"] and int(tr.stats[a, 2]) == st["found"]
            assert np.array_equal(tr.pts[a, :top], rec["pts"][:top]) and np.array_equal(tr.parent[a, :top], rec["parent"][:top])
            assert np.array_equal(tr.cost[a, :top].view(np.int64), rec["cost"][:top].view(np.int64))
            moved_any |= rec["moved"]
    assert moved_any


def test_single_agent_class_loop_runs_like_the_batch():
    """DynamicEnvironment (reference constructor, drop-in planner object) == BatchReplanner with one agent."""
    W = H = 96
    F, n, speed, rgoal = 4, 200, 4, 10.0
    og_3d = worlds.perlin_occupancygrid(W, H, frames=F, seed=3).astype(np.int64)
    planner = rrt.RRTStar(og_3d[0], n, 15.0, pbar=False, seed=11)
    env = replan.DynamicEnvironment(speed, rgoal, planner)
    goals, positions, paths, trees = env.simulate_dynamic_goals(og_3d, rnd_gen=np.random.default_rng(5))
    g = np.random.default_rng(5)
    free0 = np.argwhere(og_3d[0] == 0)
    xs, xg = free0[g.integers(low=0, high=free0.shape[0])], free0[g.integers(low=0, high=free0.shape[0])]
    rp = replan.BatchReplanner("star", W, H, n, speed, rgoal, r_rewire=15.0)
    out = rp.simulate(og_3d.astype(np.uint8), xs[None], xg[None], [11], [g])
    assert np.array_equal(out["positions"][:, 0], positions.astype(np.int64))
    assert np.array_equal(out["goals"][:, 0], goals.astype(np.int64))
    assert len(paths) == F and len(trees) == F
