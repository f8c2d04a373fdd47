"""Synthetic occupancy grids -- measurement fixture, not part of the hot path.

The reference builds its worlds with ``oggen.perlin_occupancygrid`` (``rrtplanner/oggen.py:7-45``):
fractal noise from the third-party ``pyfastnoisesimd`` library (not vendored, not installed
here), min-max normalised and thresholded so that ``noise < thresh`` is an obstacle
(``oggen.py:41-44``).  The noise arithmetic lives outside the reference tree, so its bits cannot be
matched ("parity unpinned" for the worlds -- they are inputs).  What is kept:
the function name and signature, the normalise/threshold rule and the 0 = free / 1 = obstacle
convention.

The generator below is value-noise fBm (3 octaves, lacunarity 2, gain 1/2, base frequency
about 0.02 cells^-1, i.e. blobs 40-80 cells across; about 17 % obstacle cover at the default
threshold -- SURVEY.md 8(d) expected 0.2-0.3 from the reference's figures; the cover is reported in every bench line) evaluated entirely in 32/64-bit *integer* fixed point, so that this numpy
version and the CUDA kernel ``rrtk_gen_worlds`` (csrc/grid.cu) produce bit-identical grids and
the CPU baseline and the GPU arm of ``bench.py`` see exactly the same worlds without any
cross-device copies.
"""
from __future__ import annotations

import numpy as np

FRAC_BITS = 16
ONE = 1 << FRAC_BITS
BASE_FREQ_Q16 = 1311         # 0.02 cells^-1 in Q16
OCTAVES = 3

_M1 = np.uint32(0x9E3779B1)
_M2 = np.uint32(0x85EBCA77)
_M3 = np.uint32(0xC2B2AE3D)
_M4 = np.uint32(0x2C1B3C6D)
_M5 = np.uint32(0x297A2D39)


def _lattice_hash(ix: np.ndarray, iy: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        u = ix.astype(np.uint32) * _M1 + iy.astype(np.uint32) * _M2 + np.uint32(seed & 0xFFFFFFFF) * _M3
        u ^= u >> np.uint32(15)
        u *= _M4
        u ^= u >> np.uint32(12)
        u *= _M5
        u ^= u >> np.uint32(15)
    return u


def _corner(ix, iy, seed):
    # lattice value in [0, 65535]: the top 16 bits of the hash
    return (_lattice_hash(ix, iy, seed) >> np.uint32(16)).astype(np.int64)


def _smooth(t):
    # cubic smoothstep t^2 (3 - 2 t) in Q16, all intermediates non-negative
    t2 = (t * t) >> FRAC_BITS
    return (t2 * (3 * ONE - 2 * t)) >> FRAC_BITS


def _octave(xq, yq, seed):
    ix, iy = xq >> FRAC_BITS, yq >> FRAC_BITS
    fx, fy = xq & (ONE - 1), yq & (ONE - 1)
    sx, sy = _smooth(fx), _smooth(fy)
    v00 = _corner(ix, iy, seed)
    v10 = _corner(ix + 1, iy, seed)
    v01 = _corner(ix, iy + 1, seed)
    v11 = _corner(ix + 1, iy + 1, seed)
    a = (v00 * (ONE - sx) + v10 * sx) >> FRAC_BITS
    b = (v01 * (ONE - sx) + v11 * sx) >> FRAC_BITS
    return (a * (ONE - sy) + b * sy) >> FRAC_BITS


def fbm_field(w: int, h: int, seed: int = 0) -> np.ndarray:
    """(w, h) int64 field of one world (bit-identical to the CUDA generator rrtk_gen_worlds, csrc/grid.cu)."""
    x = np.arange(w, dtype=np.int64)[:, None] * BASE_FREQ_Q16
    y = np.arange(h, dtype=np.int64)[None, :] * BASE_FREQ_Q16
    x, y = np.broadcast_arrays(x, y)
    total = np.zeros((w, h), dtype=np.int64)
    for o in range(OCTAVES):
        total += _octave(x << o, y << o, seed * 31 + o * 7919) << (OCTAVES - 1 - o)
    return total


_M6 = np.uint32(0x27D4EB2F)


def _octave3(zq, xq, yq, seed):
    """Trilinear value noise: the frame index is a third, continuous lattice coordinate, so that consecutive frames are
    neighbouring slices of one smooth 3-D field, as in the reference's genAsGrid([frames, w, h]) (oggen.py:33-38)."""
    iz, fz = zq >> FRAC_BITS, zq & (ONE - 1)
    sz = _smooth(np.int64(fz))
    with np.errstate(over="ignore"):
        s0 = int((np.uint32(seed & 0xFFFFFFFF) + np.uint32(iz & 0xFFFFFFFF) * _M6) & np.uint32(0xFFFFFFFF))
        s1 = int((np.uint32(seed & 0xFFFFFFFF) + np.uint32((iz + 1) & 0xFFFFFFFF) * _M6) & np.uint32(0xFFFFFFFF))
    return (_octave(xq, yq, s0) * (ONE - int(sz)) + _octave(xq, yq, s1) * int(sz)) >> FRAC_BITS


def fbm_field_3d(frames: int, w: int, h: int, seed: int = 0) -> np.ndarray:
    """(frames, w, h) int64 field: one frame = one step of BASE_FREQ along the third axis (the reference uses the same
    frequency on all three axes), so obstacles drift and deform from frame to frame instead of being redrawn."""
    x = np.arange(w, dtype=np.int64)[:, None] * BASE_FREQ_Q16
    y = np.arange(h, dtype=np.int64)[None, :] * BASE_FREQ_Q16
    x, y = np.broadcast_arrays(x, y)
    out = np.zeros((frames, w, h), dtype=np.int64)
    for f in range(frames):
        z = f * BASE_FREQ_Q16
        for o in range(OCTAVES):
            out[f] += _octave3(z << o, x << o, y << o, seed * 31 + o * 7919) << (OCTAVES - 1 - o)
    return out


def threshold_field(field: np.ndarray, thresh: float = 0.33) -> np.ndarray:
    """oggen.py:41-44 in exact integer form: obstacle iff (v-min)/(max-min) < thresh."""
    lo, hi = int(field.min()), int(field.max())
    num = int(round(thresh * 1000))
    return ((field - lo) * 1000 < num * (hi - lo)).astype(np.int64)


def perlin_occupancygrid(w: int, h: int, thresh: float = 0.33, frames: int = None,
                         seed: int = 0) -> np.ndarray:
    """Drop-in for ``oggen.perlin_occupancygrid`` (oggen.py:7-45) plus a ``seed``.

    Returns an int array, 1 = obstacle, 0 = free, shape (w, h) or (frames, w, h).  With
    ``frames`` the stack is a run of neighbouring slices of one 3-D noise field (a smoothly
    changing environment, which the replanning loop of anim.py:56-115 relies on) and the
    normalisation runs over the whole stack, as in the reference (oggen.py:36,41-42)."""
    if frames is None:
        return threshold_field(fbm_field(w, h, seed), thresh)
    return threshold_field(fbm_field_3d(frames, w, h, seed), thresh)


def world_seed(world_id: int) -> int:
    """Seed convention of the benchmark worlds: world w uses 1000 + w."""
    return 1000 + int(world_id)


def start_goal(og: np.ndarray, pair_id: int):
    """Start/goal pair p: two draws of free[integers(0, nfree)] from default_rng(2000 + p)
    (mirrors ``random_point_og``, rrt.py:27-44), redrawn while equal."""
    free = np.argwhere(og == 0)
    rng = np.random.default_rng(2000 + int(pair_id))
    a = free[rng.integers(0, free.shape[0])]
    b = free[rng.integers(0, free.shape[0])]
    while (a == b).all():
        b = free[rng.integers(0, free.shape[0])]
    return a.astype(np.int64), b.astype(np.int64)
