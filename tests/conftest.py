import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_plans():
    return sorted(glob.glob(os.path.join(GOLDEN, "plan_*.npz")))


def load_plan(path):
    """Golden plan record -> dict with og unpacked to a uint8 (W, H) grid."""
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    w, h = (int(v) for v in d["og_shape"])
    d["og"] = np.unpackbits(d["og"], axis=1)[:, :h].astype(np.uint8)
    assert d["og"].shape == (w, h)
    d["kind"] = str(d["kind"])
    d["n"] = int(d["n"])
    d["samples"] = d["samples"].astype(np.int64)
    d["name"] = os.path.basename(path)[5:-4]
    return d


@pytest.fixture(params=golden_plans(), ids=lambda p: os.path.basename(p)[5:-4])
def golden_plan(request):
    return load_plan(request.param)
