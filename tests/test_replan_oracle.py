"""CPU checks of oracle/replan_oracle.py (the restatement of anim.py:56-115) against known answers."""
import numpy as np

from oracle import replan_oracle as RO


def test_inflate_known_answers():
    og = np.zeros((9, 9), dtype=np.int64)
    og[4, 4] = 1
    far = RO.inflate(og, (0, 0), 1)                       # hole [0,2)x[0,2) is away from the buffer
    want = np.zeros_like(og)
    want[4, 4] = want[3, 4] = want[5, 4] = want[4, 3] = want[4, 5] = 1      # 4-connected cross (scipy default structure)
    assert np.array_equal(far, want)
    d2 = RO.inflate(og, (0, 0), 2)
    xs, ys = np.nonzero(d2)
    assert set(zip(xs, ys)) == {(x, y) for x in range(9) for y in range(9) if abs(x - 4) + abs(y - 4) <= 2} - {(x, y) for x in range(4) for y in range(4)}   # hole [0,4)x[0,4) takes (3,3)
    near = RO.inflate(og, (4, 4), 1)                      # hole [4,6)x[4,6): buffer cells (5,4), (4,5) go, the obstacle stays
    want2 = want.copy()
    want2[5, 4] = want2[4, 5] = 0
    assert np.array_equal(near, want2)
    edge = RO.inflate(og, (8, 8), 1)                      # the square is clamped onto the border cell (anim.py:84)
    assert np.array_equal(edge, want)


def test_frame_loop_moves_towards_goal_and_keeps_generator_running():
    og_3d = np.zeros((5, 40, 40), dtype=np.uint8)
    og_3d[:, 20, 5:35] = 1
    recs = RO.simulate("star", og_3d, 150, 12.0, 4, 6.0, (5, 20), (35, 20), 3, np.random.default_rng(1))
    assert len(recs) == 5
    assert any(r["moved"] for r in recs)
    assert not np.array_equal(recs[0]["samples"], recs[1]["samples"])        # one generator, successive streams (rrt.py:85)
    g = np.random.default_rng(3)
    free = np.argwhere(recs[0]["og"] == 0)
    assert np.array_equal(recs[0]["samples"], free[g.integers(0, free.shape[0], size=150)])
    for a, b in zip(recs[:-1], recs[1:]):
        step = np.abs(b["position"] - a["position"]).max()
        assert step <= 4
