"""config-5 share timing (developer tool): 125 000 synthetic datasets x 4 chains, model 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from pyhillfit_b200 import synthetic
from pyhillfit_b200.packing import SinglePack
from pyhillfit_b200.sampler import SingleLevelSampler
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
concs, Y, _ = synthetic.generate(nd)
sp = SinglePack.from_uniform(concs, Y)
ids = np.repeat(np.arange(sp.n_datasets, dtype=np.int32), 4)
for block in (0, 64, 32):
    s = SingleLevelSampler(2, sp, ids, 1.0, np.tile([6.0, 1.0, 6.0], (len(ids), 1)), variant="fit", seed=9, thinning=5, block_threads=block)
    K = 2000
    buf = torch.empty((s.n, K // 5, 4), dtype=torch.float64, device="cuda")
    s.run(K, samples=buf); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("config5 share: chains", s.n, "lanes", s.lanes, "block", s.block_threads, "stage", s.stage_groups, "rate %.3e" % (s.n * K / (best * 1e-3)))
    del buf, s
