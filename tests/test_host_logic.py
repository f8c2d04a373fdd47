"""Host-side logic that needs no GPU: seed -> PCG64 state, plan descriptors, world generator."""
import numpy as np

from rrtplanner_b200 import _lib, batch, worlds


def test_seed_states_match_numpy():
    rng = np.random.default_rng(1)
    seeds = np.concatenate([[0, 1, 2, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 32, 2 ** 63 + 11],
                            rng.integers(0, 2 ** 62, size=200), np.arange(300)]).astype(np.uint64)
    got = batch.seed_states(seeds)
    for s, g in zip(seeds, got):
        assert np.array_equal(g, _lib.pcg64_state_words(np.random.default_rng(int(s)))), int(s)


def test_make_desc_layout():
    d = batch.make_desc([3, 4], [[1, 2], [5, 6]], [[7, 8], [9, 10]], np.arange(8).reshape(2, 2, 2))
    assert d.dtype.itemsize == 64 and d["world"].tolist() == [3, 4]
    assert d["start_y"].tolist() == [2, 6] and d["goal_x"].tolist() == [7, 9]
    assert d["rot"][1].tolist() == [4, 5, 6, 7]
    raw = d.view(np.uint8).reshape(2, 64)
    assert np.frombuffer(raw[0, :20].tobytes(), dtype="<i4").tolist() == [3, 1, 2, 7, 8]


def test_world_generator_properties():
    og = worlds.perlin_occupancygrid(128, 96, seed=5)
    assert og.shape == (128, 96) and set(np.unique(og)) <= {0, 1}
    assert 0.02 < og.mean() < 0.6
    assert np.array_equal(og, worlds.perlin_occupancygrid(128, 96, seed=5))
    assert not np.array_equal(og, worlds.perlin_occupancygrid(128, 96, seed=6))
    stack = worlds.perlin_occupancygrid(64, 64, frames=3, seed=2)
    assert stack.shape == (3, 64, 64)
    a, b = worlds.start_goal(og, 3)
    assert og[a[0], a[1]] == 0 and og[b[0], b[1]] == 0 and not np.array_equal(a, b)
