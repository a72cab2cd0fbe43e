"""Hierarchical sampler: lane-per-parameter kernel vs thread-per-chain kernel over chain counts (developer tool;
run under gpurun).  usage: hier_probe.py [n_expts] [chains_per_pair,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200.packing import HierPack
from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pers = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 32, 64, 256]
K = 1000
table = Table("crumb_data")
pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne]
pr, shapes, scales, locs = hier_priors()
pack = HierPack([table.experiments(*p) for p in pairs])
print("Ne %d (%d pairs): chains, lanes, ms per %d iterations, chain-it/s, acceptance" % (ne, len(pairs), K))
for per in pers:
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), per)
    th0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(ids), 1))
    for lanes in ((16 if ne <= 5 else 32, 1, 4) if ne <= 5 else (32, 1)):
        s = HierarchicalSampler(pack, ids, th0, pr, seed=ne, thinning=5, adapt_when=100, lanes=lanes)
        buf = torch.empty((s.n, K // 5, s.d + 1), dtype=torch.float64, device="cuda")
        s.run(K, samples=buf); torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print("%7d %3d %9.2f %.3e %.3f" % (s.n, lanes, best, s.n * K / (best * 1e-3), float(s.acceptance().mean())), flush=True)
