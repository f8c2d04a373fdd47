"""K1b: the clearance-field form of RRT.collisionfree (rrt.py:183-229) -- field == brute-force Chebyshev
distance, verdicts and cell counts == the oracle's cell-by-cell walk (and the reference's golden verdicts)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from rrtplanner_b200 import _lib, batch, worlds

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def clearance_and_check(ogs, cap, segs, wid=None):
    ogs = np.ascontiguousarray(ogs, dtype=np.uint8)
    nw, W, H = ogs.shape
    db = batch.DeviceBatch("standard", W, H, 8).set_worlds_host(ogs)
    clear = torch.empty((nw, W, H), dtype=torch.uint8, device="cuda")
    scratch = torch.empty((2 * nw * db.words,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(db.L.rrtk_clearance_field(db.bits.data_ptr(), nw, W, H, cap, clear.data_ptr(), scratch.data_ptr(), st), "clearance_field")
    nseg = segs.shape[0]
    d_segs = torch.from_numpy(np.ascontiguousarray(segs, dtype=np.int32)).cuda()
    d_w = None if wid is None else torch.from_numpy(np.ascontiguousarray(wid, dtype=np.int32)).cuda()
    d_free = torch.empty(nseg, dtype=torch.uint8, device="cuda")
    d_cells = torch.empty(nseg, dtype=torch.int32, device="cuda")
    _lib.check(db.L.rrtk_collision_segments_cf(clear.data_ptr(), W, H, d_segs.data_ptr(), None if d_w is None else d_w.data_ptr(), nseg,
                                               d_free.data_ptr(), d_cells.data_ptr(), st), "collision_segments_cf")
    return clear.cpu().numpy(), d_free.cpu().numpy().astype(bool), d_cells.cpu().numpy()


def brute_clearance(og, cap):
    """min(cap, Chebyshev distance to the nearest obstacle cell or to the outside of the grid)."""
    W, H = og.shape
    pad = np.ones((W + 2 * cap, H + 2 * cap), dtype=bool)
    pad[cap:cap + W, cap:cap + H] = og != 0
    out = np.full((W, H), cap, dtype=np.int64)
    out[og != 0] = 0
    for d in range(1, cap):
        # cells whose (2d+1)^2 neighbourhood holds an obstacle have clearance <= d
        hit = np.zeros((W, H), dtype=bool)
        for dx in range(-d, d + 1):
            for dy in (-d, d):
                hit |= pad[cap + dx:cap + dx + W, cap + dy:cap + dy + H]
                hit |= pad[cap + dy:cap + dy + W, cap + dx:cap + dx + H]
        out[(out == cap) & hit] = d
    return out


@pytest.mark.parametrize("shape,cap", [((64, 64), 9), ((43, 100), 16), ((33, 31), 40), ((1, 1), 5)])
def test_field_is_capped_chebyshev_distance(shape, cap):
    rng = np.random.default_rng(shape[0] + cap)
    og = (rng.random(shape) < 0.03).astype(np.uint8)
    clear, _, _ = clearance_and_check(og[None], cap, np.zeros((1, 4), dtype=np.int32))
    assert np.array_equal(clear[0].astype(np.int64), brute_clearance(og, cap))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "collision_*.npz"))), ids=lambda p: os.path.basename(p)[10:-4])
@pytest.mark.parametrize("cap", [2, 32])
def test_golden_verdicts(path, cap):
    z = np.load(path)
    og, segs, want = z["og"], z["segs"].astype(np.int32), z["free"]
    _, free, cells = clearance_and_check((og != 0)[None], cap, segs)
    assert np.array_equal(free, want)                                   # verdicts of the unmodified reference
    assert np.array_equal(cells, c_oracle.collision_batch(og, segs)[1])   # same first-hit position


@pytest.mark.parametrize("size,nseg,cap", [(2048, 200_000, 64), (2048, 50_000, 255), (512, 100_000, 64), (97, 20_000, 7)])
def test_random_segments_vs_oracle(size, nseg, cap):
    og = worlds.perlin_occupancygrid(size, size, seed=9).astype(np.uint8)
    rng = np.random.default_rng(0)
    segs = rng.integers(0, size, size=(nseg, 4)).astype(np.int32)
    segs[:64, 2:] = segs[:64, :2]
    segs[64:128, 2] = segs[64:128, 0]
    segs[128:192, 3] = segs[128:192, 1]
    segs[192:256] = [0, 0, size - 1, size - 1]                          # the longest walk, num close to its bound
    segs[256:320] = [size - 1, 0, 0, size - 2]
    _, free, cells = clearance_and_check(og[None], cap, segs)
    want_free, want_cells = c_oracle.collision_batch(og, segs)
    assert np.array_equal(free, want_free) and np.array_equal(cells, want_cells)
    rev = segs[:20000, [2, 3, 0, 1]].copy()
    _, f2, c2 = clearance_and_check(og[None], cap, rev)
    wf, wc = c_oracle.collision_batch(og, rev)
    assert np.array_equal(f2, wf) and np.array_equal(c2, wc)


def test_empty_grid_and_ragged_block():
    og = np.zeros((300, 200), dtype=np.uint8)
    rng = np.random.default_rng(1)
    segs = np.stack([rng.integers(0, 300, 1000), rng.integers(0, 200, 1000), rng.integers(0, 300, 1000), rng.integers(0, 200, 1000)], 1)
    for nseg in (1, 31, 129, 1000):                                       # not a multiple of the per-warp block
        _, free, cells = clearance_and_check(og[None], 64, segs[:nseg])
        assert free.all()
        assert np.array_equal(cells, np.maximum(abs(segs[:nseg, 2] - segs[:nseg, 0]), abs(segs[:nseg, 3] - segs[:nseg, 1])) + 1)


def test_multi_world():
    ogs = np.stack([worlds.perlin_occupancygrid(128, 96, seed=s) for s in range(5)]).astype(np.uint8)
    rng = np.random.default_rng(4)
    nseg = 30000
    segs = np.stack([rng.integers(0, 128, nseg), rng.integers(0, 96, nseg), rng.integers(0, 128, nseg), rng.integers(0, 96, nseg)], 1).astype(np.int32)
    wid = rng.integers(0, 5, nseg).astype(np.int32)
    _, free, cells = clearance_and_check(ogs, 48, segs, wid)
    for w in range(5):
        m = wid == w
        wf, wc = c_oracle.collision_batch(ogs[w], segs[m])
        assert np.array_equal(free[m], wf) and np.array_equal(cells[m], wc)
