"""Fixed hierarchical run for ncu captures / timing (developer tool): prof_hier.py n_expts chains_per_pair iters lanes"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200.packing import HierPack
from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 3
per = int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = int(sys.argv[3]) if len(sys.argv) > 3 else 500
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
table = Table("crumb_data")
pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne]
pr, shapes, scales, locs = hier_priors()
pack = HierPack([table.experiments(*p) for p in pairs])
ids = np.repeat(np.arange(len(pairs), dtype=np.int32), per)
th0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(ids), 1))
s = HierarchicalSampler(pack, ids, th0, pr, seed=ne, thinning=5, adapt_when=100, lanes=lanes)
buf = torch.empty((s.n, K // 5, s.d + 1), dtype=torch.float64, device="cuda")
for _ in range(2):
    s.run(K, samples=buf)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("Ne", ne, "chains", s.n, "iters", K, "ms %.2f" % ms, "rate %.3e" % (s.n * K / (ms * 1e-3)), "acc %.3f" % float(s.acceptance().mean()))
