"""Small launches of every sampler kernel form for compute-sanitizer (developer tool):
    compute-sanitizer --tool memcheck  python scripts/sanitizer_run.py
    compute-sanitizer --tool racecheck python scripts/sanitizer_run.py
(the kernels' shared-memory structures: staged dose groups, draw slots / rings, gamma slots, prepared records,
covariance / factor / mean of the hierarchical thread kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200.packing import HierPack, SinglePack
from pyhillfit_b200.sampler import HierarchicalSampler, SingleLevelSampler, hier_priors, log_target_batch

table = Table("crumb_data")
pairs = table.pairs()[:6]
pack = SinglePack([table.concat(d, c) for d, c in pairs])
ids = np.repeat(np.arange(len(pairs), dtype=np.int32), 7)        # 42 chains: ragged warps
for model in (1, 2):
    d = 2 if model == 1 else 3
    th0 = np.tile([5.5, 1.0, 6.0] if model == 2 else [5.5, 6.0], (len(ids), 1))
    for lanes, spec in ((1, 1), (2, 1), (4, 1), (1, 4), (2, 4), (4, 2), (4, 8)):
        s = SingleLevelSampler(model, pack, ids, 1.0, th0, variant="temp", adapt_when=40, seed=3, thinning=5, burn_rows=4,
                               lanes=lanes, speculation=spec)
        a = s.run(93).cpu().numpy()
        b = s.run(64, row_major=True, discard_burn=True).cpu().numpy()
        assert np.isfinite(a).all() and np.isfinite(b).all()
        print("single-level model %d lanes %d speculation %d ok" % (model, lanes, spec), flush=True)
lt, _ = log_target_batch(2, pack, np.tile([5.5, 1.0, 6.0], (100, 1)), np.arange(100) % len(pairs), 1.0)
assert np.isfinite(lt.cpu().numpy()).all()
pr, shapes, scales, locs = hier_priors()
for ne in (3, 4, 6):
    hp_pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne][:3]
    hpack = HierPack([table.experiments(*p) for p in hp_pairs])
    hids = np.repeat(np.arange(len(hp_pairs), dtype=np.int32), 11)
    th0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(hids), 1))
    for lanes in ((16 if ne <= 5 else 32), 1):
        hs = HierarchicalSampler(hpack, hids, th0, pr, adapt_when=20, seed=3, thinning=5, lanes=lanes)
        a = hs.run(60).cpu().numpy()
        assert np.isfinite(a).all()
        print("hierarchical Ne %d lanes %d ok" % (ne, lanes), flush=True)
print("sanitizer run complete")
