"""Experiment: does launching the plans predicted-longest first shorten the batch?  Predictor: free cells of the world."""
import sys, numpy as np
sys.path.insert(0, '.')
import torch
import bench
from rrtplanner_b200 import _lib
P = 4096
ids = np.arange(P)
db, desc, states, _, _ = bench.cfg3_batch(0, ids)
stream = torch.cuda.current_stream()
def run(d, reps=3):
    _lib.check(d.L.rrtk_sample_streams(d.bits.data_ptr(), d.rowcum.data_ptr(), 512, 512, d.desc.data_ptr(), P, d._states.data_ptr(), 5000, d.samples.data_ptr(), stream.cuda_stream), "s")
    for _ in range(2): d.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
db._states = states
t0 = run(db)
st = db.out["stats"].cpu().numpy()
nfree = db.nfree()
acc = st[:, 9]
print("baseline order ms", round(t0, 2), " corr(nfree, accepted)", round(float(np.corrcoef(nfree, acc)[0, 1]), 3), "corr(nfree, nn_pairs)", round(float(np.corrcoef(nfree, st[:, 7])[0, 1]), 3))
for name, key in (("nfree desc", -nfree), ("accepted desc (oracle of the predictor)", -acc), ("nn_pairs desc", -st[:, 7])):
    order = np.argsort(key, kind="stable")
    d2, _, s2, _, _ = bench.cfg3_batch(0, ids[order])
    d2._states = s2
    print(name, "ms", round(run(d2), 2))
