"""Developer probe (round 2, after the prepared dose-group records freed registers): the single-level sampler at
min_ctas_hint 3 (168 registers, 3 warps per sub-partition) against 4 (128 registers, 4 warps) in the throughput
regime -- config 5's share (500 000 chains) with 1 and 2 lanes per chain -- and at config 2's size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200 import synthetic
from pyhillfit_b200.packing import SinglePack
from pyhillfit_b200.sampler import SingleLevelSampler


def rate(samplers, K):
    bufs = [torch.empty((s.n, K // 5, s.d + 1), dtype=torch.float64, device="cuda") for s in samplers]
    streams = [torch.cuda.Stream() for _ in samplers]

    def go():
        ev = torch.cuda.Event(); ev.record()
        for s, b, st in zip(samplers, bufs, streams):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                s.run(K, samples=b)
            d = torch.cuda.Event(); d.record(st)
            torch.cuda.current_stream().wait_event(d)
    go(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); go(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return sum(s.n for s in samplers) * K / (best * 1e-3)


concs, Y, _ = synthetic.generate(125000)
sp = SinglePack.from_uniform(concs, Y)
ids = np.repeat(np.arange(sp.n_datasets, dtype=np.int32), 4)
for lanes in (1, 2):
    for hint in (3, 4):
        s = SingleLevelSampler(2, sp, ids, 1.0, np.tile([6.0, 1.0, 6.0], (len(ids), 1)), variant="fit", seed=9, thinning=5, lanes=lanes)
        s.occupancy_hint = hint
        print("config-5 share, %d chains, lanes %d, hint %d: %.3e chain-it/s" % (s.n, lanes, hint, rate([s], 2000)), flush=True)
        del s

table = Table("crumb_data")
data = [table.concat(*p) for p in table.pairs()]
pack = SinglePack(data)
ids = np.repeat(np.arange(len(data), dtype=np.int32), 64)
for hint in (3, 4):
    ss = []
    for model in (1, 2):
        th = np.tile([5.5, 1.0, 6.0] if model == 2 else [5.5, 6.0], (len(ids), 1))
        s = SingleLevelSampler(model, pack, ids, 1.0, th, variant="fit", seed=3, thinning=5, lanes=2, co_resident_chains=len(ids))
        s.occupancy_hint = hint
        ss.append(s)
    print("config 2 (2 x 13 440 chains, lanes 2), hint %d: %.3e chain-it/s" % (hint, rate(ss, 10000)), flush=True)
# lanes 4 at config 2's size needs 5.7 warps per sub-partition: only hint 6 (80 registers, spills) holds it -- not run
