"""CPU-side checks of the boundary: the C-ABI library builds/loads, exports every symbol the header
declares, and fails loudly (never falls back) when no GPU is present.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rrtplanner_b200 import _lib, build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_mod.build()
    return _lib.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rrtk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rrtk_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rrtk.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_layout(lib):
    assert lib.rrtk_version() == 100
    assert _lib.PLAN_DESC.itemsize == 64
    assert lib.rrtk_grid_words(512, 512) == 512 * 512 // 32
    assert lib.rrtk_grid_words(43, 100) == 2 * 4 * 32      # padded to whole 32x32 tiles
    assert lib.rrtk_grid_words(0, 5) == 0


def test_invalid_arguments_are_rejected_without_a_gpu(lib):
    assert lib.rrtk_pack_grid(None, 1, 8, 8, None, None) == -1
    assert b"null" in lib.rrtk_last_error()
    assert lib.rrtk_plan_batch(1, None, 8, 8, None, 1, 10, 1.0, 1.0, None, None, None, None, None, None, None, 0, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(MemoryError):
        _lib.check(-2, "x")


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device the product path must raise, not compute on the host."""
    import rrtplanner_b200 as R
    if lib.rrtk_device_count() > 0:
        pytest.skip("GPU present")
    og = np.zeros((16, 16))
    p = R.RRTStar(og, 10, 5, pbar=False)
    with pytest.raises(RuntimeError):
        p.plan(np.array([1, 1]), np.array([9, 9]))
    with pytest.raises(RuntimeError):
        R.RRT.collisionfree(og, np.array([1, 1]), np.array([2, 2]))
    with pytest.raises(NotImplementedError):
        R.RRT(og, 10).plan(np.array([1, 1]), np.array([2, 2]))
    with pytest.raises(NotImplementedError):
        R.RRTStar(og, 10, 5, costfn=lambda *a: 0.0)
    # the planners the reference only advertises behave the same way: no GPU, no result
    with pytest.raises(RuntimeError):
        R.RRTStar(og, 10, 5, pbar=False, rewire="rrtstar").plan(np.array([1, 1]), np.array([9, 9]))
    with pytest.raises(RuntimeError):
        R.RRTStarDubins(og, 10, 5.0, 2.0, pbar=False).plan(np.array([1, 1, 0]), np.array([9, 9, 3]))
    with pytest.raises(RuntimeError):
        R.dubins_path([1, 1, 0], [9, 9, 3], 2.0)


def test_k8_argument_checks_need_no_gpu(lib):
    cfg = _lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 5.0, 16, 2.0, 1.0)
    assert lib.rrtk_plan2_batch(_lib.ptr(cfg), None, 8, 8, None, 1, 10, None, None, None, None, None, None, None, None, None, 0, None) == -1
    bad = _lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 5.0, 0, 2.0, 1.0)           # zero headings
    assert lib.rrtk_plan2_batch(_lib.ptr(bad), None, 8, 8, None, 1, 10, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert b"nheadings" in lib.rrtk_last_error()
    tiny = _lib.plan2_cfg(_lib.MODEL_DUBINS, True, True, 5.0, 16, 2.0, 1e-9)        # a step that would sample ~1e10 points per path
    assert lib.rrtk_plan2_batch(_lib.ptr(tiny), None, 8, 8, None, 1, 10, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.rrtk_plan2_scratch_bytes(4, 100) >= 4 * 101 * 10 and lib.rrtk_plan2_scratch_bytes(-1, 100) == 0
    assert lib.rrtk_dubins_table_bytes(50, 16) == (101 * 101 * 256) * 33 + 16
    assert lib.rrtk_dubins_paths(None, 5, 16, 2.0, None, None, None, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rrtplanner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_pcg_state_words():
    w = _lib.pcg64_state_words(np.random.default_rng(0))
    st = np.random.default_rng(0).bit_generator.state["state"]
    assert (int(w[0]) << 64) | int(w[1]) == st["state"] and (int(w[2]) << 64) | int(w[3]) == st["inc"]


def test_struct_layouts_match_the_header(tmp_path):
    """The numpy mirrors of the ABI structs (rrtk_plan_desc, rrtk_plan2_cfg) have the sizes and field offsets a C compiler
    gives the declarations of include/rrtk.h."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "rrtk.h"
#define F(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
    printf("rrtk_plan_desc %zu\n", sizeof(rrtk_plan_desc));
    F(rrtk_plan_desc, world); F(rrtk_plan_desc, start_x); F(rrtk_plan_desc, start_y); F(rrtk_plan_desc, goal_x);
    F(rrtk_plan_desc, goal_y); F(rrtk_plan_desc, reserved); F(rrtk_plan_desc, rot);
    printf("rrtk_plan2_cfg %zu\n", sizeof(rrtk_plan2_cfg));
    F(rrtk_plan2_cfg, model); F(rrtk_plan2_cfg, star); F(rrtk_plan2_cfg, rewire); F(rrtk_plan2_cfg, nheadings);
    F(rrtk_plan2_cfg, r_rewire); F(rrtk_plan2_cfg, rho); F(rrtk_plan2_cfg, ds); F(rrtk_plan2_cfg, dubins_table);
    F(rrtk_plan2_cfg, table_radius); F(rrtk_plan2_cfg, informed); F(rrtk_plan2_cfg, r_goal); F(rrtk_plan2_cfg, balls);
    F(rrtk_plan2_cfg, ell_c);
    printf("RRTK_STAT_COUNT %d\n", (int)RRTK_STAT_COUNT);
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.rsplit(" ", 1) for line in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    assert int(got["rrtk_plan_desc"]) == _lib.PLAN_DESC.itemsize and int(got["rrtk_plan2_cfg"]) == _lib.PLAN2_CFG.itemsize
    for struct, dt in (("rrtk_plan_desc", _lib.PLAN_DESC), ("rrtk_plan2_cfg", _lib.PLAN2_CFG)):
        for name in dt.names:
            assert int(got[f"{struct}.{name}"]) == dt.fields[name][1], (struct, name)
    assert int(got["RRTK_STAT_COUNT"]) == _lib.STAT_COUNT == len(_lib.STAT2_NAMES)
