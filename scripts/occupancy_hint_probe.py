"""Developer probe: register budget of the single-level sampler (cfg.min_ctas_hint: 3 -> 168 registers, 4 -> 128) at the
throughput end (one thread per chain, 215 040 chains) and at config 2's occupancy (2 lanes, 26 880 chains of model 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from bench import build_workload
from pyhillfit_b200.sampler import SingleLevelSampler
K = 2000
for cpp, lanes in ((1024, 1), (2048, 1), (128, 2), (256, 2)):
    pack, wl = build_workload(cpp)
    w = wl[2]
    for hint in (3, 4):
        s = SingleLevelSampler(2, pack, w["ids"], 1.0, w["theta0"], variant="fit", seed=25, thinning=5, lanes=lanes)
        s.occupancy_hint = hint
        buf = torch.empty((K // 5 + 1, s.n, 4), dtype=torch.float64, device="cuda")
        s.run(K, samples=buf, row_major=True); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.run(K, samples=buf, row_major=True); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print("chains %7d lanes %d min_ctas_hint %d: %8.2f ms per %d iterations  %.3e chain-it/s" % (s.n, lanes, hint, best, K, s.n * K / (best * 1e-3)), flush=True)
        del buf, s
