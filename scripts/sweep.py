"""Throughput sweep of the single-level sampler (developer tool; run under gpurun)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from bench import build_workload
from pyhillfit_b200.sampler import SingleLevelSampler

def rate(model, cpp, block, stage, K=4000, reps=3, two=False):
    pack, wl = build_workload(cpp)
    w = wl[model]
    s = SingleLevelSampler(model, pack, w["ids"], 1.0, w["theta0"], variant="fit", seed=25, thinning=5,
                           stage=stage, block_threads=block)
    buf = torch.empty((s.n, K // 5 + 1, w["d"] + 1), dtype=torch.float64, device="cuda")
    s.run(K, samples=buf); s.run(K, samples=buf)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return s.n * K / (best * 1e-3), best, s.stage_groups, s.block_threads

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("model cpp block stage -> rate (chain-it/s), ms, stage_groups")
    for model in (2, 1):
        for cpp in (64, 128, 256, 512, 1024, 2048):
            for block in ((0, 32, 64, 128) if cpp in (64, 512) else (0,)):
                r, ms, sg, bt = rate(model, cpp, block, True)
                print(model, cpp, bt, "stage", "%.3e" % r, "%.2f" % ms, sg, flush=True)
        r, ms, sg, bt = rate(model, 64, 32, False)
        print(model, 64, bt, "nostage", "%.3e" % r, "%.2f" % ms, sg, flush=True)
