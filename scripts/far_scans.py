import sys, numpy as np
sys.path.insert(0, '.')
import torch, bench
from rrtplanner_b200 import _lib
P = 1036
db, desc, states, _, _ = bench.cfg3_batch(0, np.arange(P))
db.seed_samples(np.arange(P)); db.run(); torch.cuda.synchronize()
st = db.out["stats"].cpu().numpy()
names = list(_lib.STAT_NAMES)
far = st[:, names.index("reserved0")]; acc = st[:, names.index("accepted")]; j = st[:, 0]
print("far scans per plan: mean %.0f p50 %.0f p90 %.0f max %d; accepted mean %.0f" % (far.mean(), np.percentile(far, 50), np.percentile(far, 90), far.max(), acc.mean()))
