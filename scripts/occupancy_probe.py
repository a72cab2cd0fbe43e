"""Per-iteration latency of the single-level sampler as a function of resident warps (developer tool; run under
gpurun).  One dataset (Amiodarone/hERG), 32-thread CTAs so that the number of warps per SM is known:
chains = warps_per_sm * 148 * 32 / lanes.   usage: occupancy_probe.py [model] [lanes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200.packing import SinglePack
from pyhillfit_b200.sampler import SingleLevelSampler

model = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
table = Table("crumb_data")
pack = SinglePack([table.concat("Amiodarone", "hERG")])
sms = torch.cuda.get_device_properties(0).multi_processor_count
K = 4000
print("model %d lanes %d: warps/SM, chains, ms, cycles per iteration at 1965 MHz, chain-it/s" % (model, lanes))
for wps in (0, 1, 2, 4, 6, 8, 10, 12):
    n = 32 // lanes if wps == 0 else wps * sms * 32 // lanes
    th = np.tile([5.5, 1.0, 6.0] if model == 2 else [5.5, 6.0], (n, 1)) * (1 + 0.02 * np.random.default_rng(1).standard_normal((n, 3 if model == 2 else 2)))
    s = SingleLevelSampler(model, pack, np.zeros(n, np.int32), 1.0, th, variant="fit", seed=3, thinning=5,
                           block_threads=32, lanes=lanes)
    buf = torch.empty((n, K // 5 + 1, th.shape[1] + 1), dtype=torch.float64, device="cuda")
    s.run(K, samples=buf); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("%4s %7d %8.3f %8.0f %.3e" % (wps if wps else "1w", n, best, best * 1e-3 / K * 1.965e9, n * K / (best * 1e-3)), flush=True)
