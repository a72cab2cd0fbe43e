"""Throughput sweep of the single-level sampler (developer tool; run under gpurun).
usage: sweep.py [model,...] [chains_per_pair,...] [lanes:minb,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from bench import build_workload
from pyhillfit_b200.sampler import SingleLevelSampler

_wl = {}
def rate(model, cpp, lanes, minb, block=0, K=4000, reps=3):
    if cpp not in _wl:
        _wl[cpp] = build_workload(cpp)
    pack, wl = _wl[cpp]
    w = wl[model]
    s = SingleLevelSampler(model, pack, w["ids"], 1.0, w["theta0"], variant="fit", seed=25, thinning=5,
                           block_threads=block, lanes=lanes)
    s.occupancy_hint = minb
    buf = torch.empty((s.n, K // 5 + 1, w["d"] + 1), dtype=torch.float64, device="cuda")
    s.run(K, samples=buf); s.run(K, samples=buf)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.run(K, samples=buf); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return s.n * K / (best * 1e-3), best, s.stage_groups, s.block_threads, float(s.acceptance().mean())

if __name__ == "__main__":
    models = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2]
    cpps = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 128, 256, 1024]
    variants = [tuple(int(y) for y in x.split(":")) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [(1, 0), (2, 0), (4, 0)]
    print("model chains_per_pair lanes minb block -> rate (chain-it/s), ms, stage_groups, acceptance")
    for model in models:
        for cpp in cpps:
            for v in variants:
                lanes, minb = v[0], v[1]
                block = v[2] if len(v) > 2 else 0
                try:
                    r, ms, sg, bt, acc = rate(model, cpp, lanes, minb, block)
                    print(model, cpp, lanes, minb, bt, "%.3e" % r, "%.2f" % ms, sg, "%.3f" % acc, flush=True)
                except Exception as e:
                    print(model, cpp, lanes, minb, "ERR", e, flush=True)
