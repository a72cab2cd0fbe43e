"""Small fixed run for ncu captures (developer tool): single-level sampler at a chosen chain count.
usage: prof_run.py chains_per_pair iters models lanes minb speculation"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from bench import build_workload
from pyhillfit_b200.sampler import SingleLevelSampler
cpp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 500
models = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2]
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
minb = int(sys.argv[5]) if len(sys.argv) > 5 else 0
spec = int(sys.argv[6]) if len(sys.argv) > 6 else 1
pack, wl = build_workload(cpp)
for model in models:
    w = wl[model]
    s = SingleLevelSampler(model, pack, w["ids"], 1.0, w["theta0"], variant="fit", seed=25, thinning=5, adapt_when=100,
                           lanes=lanes, speculation=spec)
    s.occupancy_hint = minb
    buf = torch.empty((s.n, K // 5 + 1, w["d"] + 1), dtype=torch.float64, device="cuda")
    for _ in range(3):
        s.run(K, samples=buf)
    torch.cuda.synchronize()
    print("model", model, "chains", s.n, "lanes", s.lanes, "speculation", s.speculation, "acc", float(s.acceptance().mean()))
