import os, sys
ROOT = "/root/repo" if os.path.isdir("/root/repo/tests") else os.getcwd()
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200.packing import HierPack
from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
table = Table("crumb_data"); pr, *_ = hier_priors(); K = 1000
for ne, pers in ((3, (64, 128, 256)), (4, (128, 256)), (5, (256,))):
    pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne]
    pack = HierPack([table.experiments(*p) for p in pairs])
    for per in pers:
        ids = np.repeat(np.arange(len(pairs), dtype=np.int32), per)
        th0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(ids), 1))
        for lanes, hint, bt in ((1, 0, 0), (4, 3, 0), (4, 4, 0), (4, 3, 64), (4, 4, 64)):
            s = HierarchicalSampler(pack, ids, th0, pr, seed=ne, thinning=5, adapt_when=100, lanes=lanes, block_threads=bt if lanes == 4 else 0)
            s.occupancy_hint = hint
            buf = torch.empty((s.n, K // 5, s.d + 1), dtype=torch.float64, device="cuda")
            s.run(K, samples=buf); torch.cuda.synchronize()
            best = 1e9
            for _ in range(2):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            print("Ne %d chains %6d lanes %d regs-hint %d block %3d: %7.2f ms  %.3e" % (ne, s.n, lanes, hint, bt, best, s.n * K / (best * 1e-3)), flush=True)
