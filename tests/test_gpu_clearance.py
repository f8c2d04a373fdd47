"""K1b: the clearance-field forms of RRT.collisionfree (rrt.py:183-229) -- isotropic field == brute-force Chebyshev
distance, directional fields == the depth of the free cone ahead of a cell, verdicts and cell counts of both == the
oracle's cell-by-cell walk (and the reference's golden verdicts)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from rrtplanner_b200 import _lib, batch, worlds

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


MODES = ["iso", "dir", "dir16"]


def clearance_and_check(ogs, cap, segs, wid=None, mode="iso"):
    ogs = np.ascontiguousarray(ogs, dtype=np.uint8)
    nw, W, H = ogs.shape
    db = batch.DeviceBatch("standard", W, H, 8).set_worlds_host(ogs)
    st = torch.cuda.current_stream().cuda_stream
    if mode == "iso":
        clear = torch.empty((nw, W, H), dtype=torch.uint8, device="cuda")
        scratch = torch.empty((2 * nw * db.words,), dtype=torch.int32, device="cuda")
        _lib.check(db.L.rrtk_clearance_field(db.bits.data_ptr(), nw, W, H, cap, clear.data_ptr(), scratch.data_ptr(), st), "clearance_field")
        walk = db.L.rrtk_collision_segments_cf
    elif mode == "dir":
        clear = torch.full((nw, 8, W, H), 77, dtype=torch.uint8, device="cuda")
        _lib.check(db.L.rrtk_clearance_field_dir(db.bits.data_ptr(), nw, W, H, cap, clear.data_ptr(), st), "clearance_field_dir")
        walk = db.L.rrtk_collision_segments_cfd
    else:
        clear = torch.full((nw, 16, W, H), 77, dtype=torch.uint8, device="cuda")
        _lib.check(db.L.rrtk_clearance_field_dir16(db.bits.data_ptr(), nw, W, H, cap, clear.data_ptr(), st), "clearance_field_dir16")
        walk = db.L.rrtk_collision_segments_cfd16
    nseg = segs.shape[0]
    d_segs = torch.from_numpy(np.ascontiguousarray(segs, dtype=np.int32)).cuda()
    d_w = None if wid is None else torch.from_numpy(np.ascontiguousarray(wid, dtype=np.int32)).cuda()
    d_free = torch.empty(nseg, dtype=torch.uint8, device="cuda")
    d_cells = torch.empty(nseg, dtype=torch.int32, device="cuda")
    _lib.check(walk(clear.data_ptr(), W, H, d_segs.data_ptr(), None if d_w is None else d_w.data_ptr(), nseg,
                    d_free.data_ptr(), d_cells.data_ptr(), st), "collision_segments_cf(d)")
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


def cone_depth(og, cap, octant):
    """Directional field of one octant, by definition: D = 0 on obstacles, else min(cap, 1 + min(D(one step ahead on the major
    axis), D(one step ahead on both axes))); outside the grid counts as free.  octant = 4 * xmajor + 2 * (dx > 0) + (dy > 0)."""
    xmajor, xpos, ypos = bool(octant & 4), bool(octant & 2), bool(octant & 1)
    occ = (og != 0) if xmajor else (og != 0).T                     # axis 0 = major
    smaj, smin = (xpos, ypos) if xmajor else (ypos, xpos)
    occ = occ[::1 if smaj else -1, ::1 if smin else -1]            # the walk now goes towards +axis0, +axis1
    n0, n1 = occ.shape
    D = np.full((n0 + 1, n1 + 1), cap, dtype=np.int64)
    for a in range(n0 - 1, -1, -1):
        D[a, :n1] = np.where(occ[a], 0, np.minimum(cap, 1 + np.minimum(D[a + 1, :n1], D[a + 1, 1:])))
    D = D[:n0, :n1][::1 if smaj else -1, ::1 if smin else -1]
    return D if xmajor else D.T


@pytest.mark.parametrize("shape,cap", [((64, 64), 9), ((43, 100), 16), ((33, 31), 40), ((1, 1), 5), ((1300, 70), 255), ((50, 2100), 200)])
def test_directional_fields_are_capped_cone_depths(shape, cap):
    rng = np.random.default_rng(shape[0] + cap)
    og = (rng.random(shape) < 0.01).astype(np.uint8)
    clear, _, _ = clearance_and_check(og[None], cap, np.zeros((1, 4), dtype=np.int32), mode="dir")
    for octant in range(8):
        assert np.array_equal(clear[0, octant].astype(np.int64), cone_depth(og, cap, octant)), octant


def half_cone_depths(og, cap, octant):
    """The two fields of an octant split at slope 1/2, by definition (include/rrtk.h): low half A = 1 + min(B(+a), B(+a+b)),
    B = 1 + A(+a); high half C = 1 + min(D(+a), D(+a+b)), D = 1 + C(+a+b); 0 on obstacles, capped, outside the grid free."""
    xmajor, xpos, ypos = bool(octant & 4), bool(octant & 2), bool(octant & 1)
    occ = (og != 0) if xmajor else (og != 0).T
    smaj, smin = (xpos, ypos) if xmajor else (ypos, xpos)
    occ = occ[::1 if smaj else -1, ::1 if smin else -1]
    n0, n1 = occ.shape
    A = np.full((n0 + 1, n1 + 1), cap, dtype=np.int64); B = A.copy(); C = A.copy(); D = A.copy()
    for a in range(n0 - 1, -1, -1):
        o = occ[a]
        A[a, :n1] = np.where(o, 0, np.minimum(cap, 1 + np.minimum(B[a + 1, :n1], B[a + 1, 1:])))
        B[a, :n1] = np.where(o, 0, np.minimum(cap, 1 + A[a + 1, :n1]))
        C[a, :n1] = np.where(o, 0, np.minimum(cap, 1 + np.minimum(D[a + 1, :n1], D[a + 1, 1:])))
        D[a, :n1] = np.where(o, 0, np.minimum(cap, 1 + C[a + 1, 1:]))
    out = []
    for F in (A, C):
        F = F[:n0, :n1][::1 if smaj else -1, ::1 if smin else -1]
        out.append(F if xmajor else F.T)
    return out


@pytest.mark.parametrize("shape,cap", [((64, 64), 9), ((43, 100), 16), ((33, 31), 40), ((1, 1), 5), ((1300, 70), 255), ((50, 2100), 200)])
def test_half_octant_fields_are_capped_cone_depths(shape, cap):
    rng = np.random.default_rng(shape[0] + cap + 1)
    og = (rng.random(shape) < 0.01).astype(np.uint8)
    clear, _, _ = clearance_and_check(og[None], cap, np.zeros((1, 4), dtype=np.int32), mode="dir16")
    for octant in range(8):
        low, high = half_cone_depths(og, cap, octant)
        assert np.array_equal(clear[0, 2 * octant].astype(np.int64), low), octant
        assert np.array_equal(clear[0, 2 * octant + 1].astype(np.int64), high), octant


def test_half_octant_fields_dominate_the_octant_fields():
    og = worlds.perlin_occupancygrid(200, 168, seed=3).astype(np.uint8)
    d8, _, _ = clearance_and_check(og[None], 64, np.zeros((1, 4), dtype=np.int32), mode="dir")
    d16, _, _ = clearance_and_check(og[None], 64, np.zeros((1, 4), dtype=np.int32), mode="dir16")
    for octant in range(8):
        assert (d16[0, 2 * octant] >= d8[0, octant]).all() and (d16[0, 2 * octant + 1] >= d8[0, octant]).all()


def test_directional_fields_dominate_the_isotropic_field():
    og = worlds.perlin_occupancygrid(200, 168, seed=3).astype(np.uint8)
    iso, _, _ = clearance_and_check(og[None], 64, np.zeros((1, 4), dtype=np.int32))
    dirs, _, _ = clearance_and_check(og[None], 64, np.zeros((1, 4), dtype=np.int32), mode="dir")
    assert (dirs[0] >= iso[0][None]).all() and ((dirs[0] == 0) == (og != 0)[None]).all()


@pytest.mark.parametrize("shape,cap", [((64, 64), 9), ((43, 100), 16), ((33, 31), 40), ((1, 1), 5)])
def test_field_is_capped_chebyshev_distance(shape, cap):
    rng = np.random.default_rng(shape[0] + cap)
    og = (rng.random(shape) < 0.03).astype(np.uint8)
    clear, _, _ = clearance_and_check(og[None], cap, np.zeros((1, 4), dtype=np.int32))
    assert np.array_equal(clear[0].astype(np.int64), brute_clearance(og, cap))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "collision_*.npz"))), ids=lambda p: os.path.basename(p)[10:-4])
@pytest.mark.parametrize("cap", [2, 32])
@pytest.mark.parametrize("mode", MODES)
def test_golden_verdicts(path, cap, mode):
    z = np.load(path)
    og, segs, want = z["og"], z["segs"].astype(np.int32), z["free"]
    _, free, cells = clearance_and_check((og != 0)[None], cap, segs, mode=mode)
    assert np.array_equal(free, want)                                   # verdicts of the unmodified reference
    assert np.array_equal(cells, c_oracle.collision_batch(og, segs)[1])   # same first-hit position


@pytest.mark.parametrize("size,nseg,cap", [(2048, 200_000, 64), (2048, 50_000, 255), (512, 100_000, 64), (97, 20_000, 7)])
@pytest.mark.parametrize("mode", MODES)
def test_random_segments_vs_oracle(size, nseg, cap, mode):
    og = worlds.perlin_occupancygrid(size, size, seed=9).astype(np.uint8)
    rng = np.random.default_rng(0)
    segs = rng.integers(0, size, size=(nseg, 4)).astype(np.int32)
    segs[:64, 2:] = segs[:64, :2]
    segs[64:128, 2] = segs[64:128, 0]
    segs[128:192, 3] = segs[128:192, 1]
    segs[192:256] = [0, 0, size - 1, size - 1]                          # the longest walk, num close to its bound
    segs[256:320] = [size - 1, 0, 0, size - 2]
    _, free, cells = clearance_and_check(og[None], cap, segs, mode=mode)
    want_free, want_cells = c_oracle.collision_batch(og, segs)
    assert np.array_equal(free, want_free) and np.array_equal(cells, want_cells)
    rev = segs[:20000, [2, 3, 0, 1]].copy()
    _, f2, c2 = clearance_and_check(og[None], cap, rev, mode=mode)
    wf, wc = c_oracle.collision_batch(og, rev)
    assert np.array_equal(f2, wf) and np.array_equal(c2, wc)


@pytest.mark.parametrize("mode", MODES)
def test_empty_grid_and_ragged_block(mode):
    og = np.zeros((300, 200), dtype=np.uint8)
    rng = np.random.default_rng(1)
    segs = np.stack([rng.integers(0, 300, 1000), rng.integers(0, 200, 1000), rng.integers(0, 300, 1000), rng.integers(0, 200, 1000)], 1)
    for nseg in (1, 31, 129, 1000):                                       # not a multiple of the per-warp block
        _, free, cells = clearance_and_check(og[None], 64, segs[:nseg], mode=mode)
        assert free.all()
        assert np.array_equal(cells, np.maximum(abs(segs[:nseg, 2] - segs[:nseg, 0]), abs(segs[:nseg, 3] - segs[:nseg, 1])) + 1)


@pytest.mark.parametrize("mode", MODES)
def test_multi_world(mode):
    ogs = np.stack([worlds.perlin_occupancygrid(128, 96, seed=s) for s in range(5)]).astype(np.uint8)
    rng = np.random.default_rng(4)
    nseg = 30000
    segs = np.stack([rng.integers(0, 128, nseg), rng.integers(0, 96, nseg), rng.integers(0, 128, nseg), rng.integers(0, 96, nseg)], 1).astype(np.int32)
    wid = rng.integers(0, 5, nseg).astype(np.int32)
    _, free, cells = clearance_and_check(ogs, 48, segs, wid, mode=mode)
    for w in range(5):
        m = wid == w
        wf, wc = c_oracle.collision_batch(ogs[w], segs[m])
        assert np.array_equal(free[m], wf) and np.array_equal(cells[m], wc)
