"""Config 2 (models 1 and 2 co-resident on two streams): step time for every (lanes model 1, lanes model 2, CTA size)
combination (developer tool; run under gpurun).  usage: lane_combo.py [chains_per_pair]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from bench import build_workload
from pyhillfit_b200.sampler import SingleLevelSampler

cpp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pack, wl = build_workload(cpp)
K = 5000
streams = {m: torch.cuda.Stream() for m in (1, 2)}
combos = [(2, 2, 0, 0), (1, 2, 0, 0), (2, 1, 0, 0), (1, 1, 0, 0), (1, 2, 32, 64), (1, 2, 32, 32), (1, 2, 64, 32), (4, 2, 0, 0), (1, 4, 0, 0),
          (1, 1, 32, 32), (2, 2, 32, 32), (2, 2, 128, 128)]
print("lanes(m1) lanes(m2) block(m1) block(m2) -> ms per %d iterations, chain-it/s" % K)
for l1, l2, b1, b2 in combos:
    S, B = {}, {}
    for m, l, b in ((1, l1, b1), (2, l2, b2)):
        w = wl[m]
        S[m] = SingleLevelSampler(m, pack, w["ids"], 1.0, w["theta0"], variant="fit", seed=25, thinning=5, block_threads=b, lanes=l)
        B[m] = torch.empty((S[m].n, K // 5 + 1, w["d"] + 1), dtype=torch.float64, device="cuda")
    def step():
        ev = torch.cuda.Event(); ev.record()
        for m in (1, 2):
            streams[m].wait_event(ev)
            with torch.cuda.stream(streams[m]):
                S[m].run(K, samples=B[m])
            done = torch.cuda.Event(); done.record(streams[m]); torch.cuda.current_stream().wait_event(done)
    step(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    n = S[1].n + S[2].n
    print(l1, l2, S[1].block_threads, S[2].block_threads, "%.3f" % best, "%.3e" % (n * K / (best * 1e-3)), flush=True)
    del S, B
