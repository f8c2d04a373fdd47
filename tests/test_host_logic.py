"""Host-side logic that needs no GPU: seed -> PCG64 state, plan descriptors, world generator."""
import numpy as np

from rrtplanner_b200 import _lib, batch, worlds


def test_seed_states_match_numpy():
    rng = np.random.default_rng(1)
    seeds = np.concatenate([[0, 1, 2, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 32, 2 ** 63 + 11],
                            rng.integers(0, 2 ** 62, size=200), np.arange(300)]).astype(np.uint64)
    got = batch.seed_states(seeds)                       # rrtk_seed_states (C, host)
    assert np.array_equal(got, batch.seed_states_numpy(seeds))
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
    # frames are neighbouring slices of one smooth 3-D field (oggen.py:33-38): the environment changes gradually
    long = worlds.perlin_occupancygrid(128, 128, frames=8, seed=2)
    step = [(long[i] != long[i + 1]).mean() for i in range(7)]
    assert 0 < max(step) < 0.05 and (long[0] != long[7]).mean() > max(step)
    a, b = worlds.start_goal(og, 3)
    assert og[a[0], a[1]] == 0 and og[b[0], b[1]] == 0 and not np.array_equal(a, b)


def test_overridden_primitives_are_refused_not_ignored():
    """rrt.py:131,157,183 are the reference's override points; plan() is one fused kernel, so a subclass that
    replaces one of them must get an error (like costfn), never the stock behaviour.  Needs no GPU: the check runs first."""
    import pytest
    from rrtplanner_b200 import rrt

    og = np.zeros((16, 16), dtype=np.uint8)
    for name in ("near", "within", "collisionfree"):
        for base, args in ((rrt.RRTStandard, (og, 10)), (rrt.RRTStar, (og, 10, 5.0)), (rrt.RRTStarInformed, (og, 10, 5.0, 2.0))):
            Sub = type("Sub", (base,), {name: staticmethod(lambda *a, **k: None)})
            with pytest.raises(NotImplementedError, match=name):
                Sub(*args, pbar=False).plan(np.array([1, 1]), np.array([5, 5]))

    class Mine(rrt.RRTStar):
        def sample_all_free(self):
            return np.array([3, 4])

    assert Mine(og, 10, 5.0, pbar=False)._sampler_overridden() and not rrt.RRTStar(og, 10, 5.0, pbar=False)._sampler_overridden()
    got = Mine(og, 4, 5.0, pbar=False)._draw_samples(4)
    assert got.dtype == np.int64 and got.tolist() == [[3, 4]] * 4
    from rrtplanner_b200 import dubins
    assert not dubins.RRTDubins(og, 4, 2.0, pbar=False)._sampler_overridden()       # its own sampler is the stock one

    class Bad(rrt.RRTStandard):
        def sample_all_free(self):
            return np.array([3, 99])

    with pytest.raises(ValueError, match="outside"):
        Bad(og, 4, pbar=False)._draw_samples(2)


def test_plan_batch_informed_needs_streams():
    import pytest
    og = np.zeros((1, 32, 32), dtype=np.uint8)
    with pytest.raises(ValueError, match="balls"):
        batch.plan_batch("informed", og, 50, [[1, 1]], [[20, 20]], r_rewire=5.0, r_goal=2.0, seeds=[0])


def test_host_packer_matches_the_documented_bit_layout():
    """rrtk_pack_grid_host (K0 on the CPU, for callers that keep worlds packed) against the numpy statement of the tiled
    layout documented in include/rrtk.h; the GPU packer is held to the same statement in tests/test_gpu_parity.py."""
    rng = np.random.default_rng(4)
    for W, H in ((64, 64), (43, 100), (100, 43), (33, 31), (1, 1), (96, 160)):
        og = (rng.random((2, W, H)) < 0.3) * rng.integers(1, 200, size=(2, W, H))
        got = _lib.pack_grids_host(og)
        TX, TY = (W + 31) // 32, (H + 31) // 32
        assert got.shape == (2, TX * TY * 32) and got.dtype == np.uint32
        for w in range(2):
            pad = np.ones((TX * 32, TY * 32), dtype=np.uint64)
            pad[:W, :H] = og[w] != 0
            t = pad.reshape(TX, 32, TY, 32).transpose(0, 2, 1, 3)
            want = (t << np.arange(32, dtype=np.uint64)).sum(axis=3).reshape(-1).astype(np.uint32)
            assert np.array_equal(got[w], want)
