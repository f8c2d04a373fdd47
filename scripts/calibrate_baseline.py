"""CPU container only (needs /root/reference): time the UNMODIFIED reference planner (rrtplanner/rrt.py loaded by file
path) and the oracle's port (oracle/rrt_oracle.py:plan_star, what bench.py's CPU arm runs on the GPU box, where the
reference cannot travel) on the same plans of the cfg3 workload, one core.  The ratio is recorded in BASELINE.md."""
import importlib.util
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, ".")
import bench                                            # noqa: E402
from oracle import rrt_oracle as O                      # noqa: E402

spec = importlib.util.spec_from_file_location("ref_rrt", "/root/reference/rrtplanner/rrt.py")
ref = importlib.util.module_from_spec(spec)
warnings.simplefilter("ignore")
spec.loader.exec_module(ref)

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 3
bench._cpu_warm()
og0 = np.zeros((32, 32), dtype=np.int64)
ref.RRTStar(og0, 20, 5, pbar=False).plan(np.array([1, 1]), np.array([20, 20]))          # JIT warm-up of the reference's two Numba functions
t_ref = t_port = 0.0
same = True
for pid in range(NP):
    og, xs, xg = bench.host_world_and_pair(pid)
    smp = O.sample_stream(og, bench.N_ITER, pid)
    t0 = time.perf_counter()
    tree = O.plan_star(og, bench.N_ITER, bench.R_REWIRE, xs, xg, smp)
    t_port += time.perf_counter() - t0
    pl = ref.RRTStar(og.astype(np.int64), bench.N_ITER, bench.R_REWIRE, pbar=False, seed=pid)   # same stream: default_rng(pid).integers(0, nfree, n)
    t0 = time.perf_counter()
    T, gv = pl.plan(xs, xg)
    t_ref += time.perf_counter() - t0
    # the unmodified reference breaks nearest ties with an unstable sort, so trees may differ from the pinned port in a few vertices;
    # vertex counts are reported, not asserted
    same = same and (T.number_of_nodes() in (bench.N_ITER, bench.N_ITER + 1))
print(f"plans {NP}: unmodified reference {t_ref / NP:.2f} s/plan ({NP / t_ref:.3f} plans/s/core), port {t_port / NP:.2f} s/plan "
      f"({NP / t_port:.3f} plans/s/core), reference / port time ratio {t_ref / t_port:.2f}")
