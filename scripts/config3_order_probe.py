"""Developer probe: BASELINE config 3's four launches (library's kernel choice, each told what runs beside it) under
different launch orders and stream priorities.  Prints ms per 1000 iterations for the concurrent run."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200.packing import HierPack
from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors

table = Table("crumb_data")
pr, shapes, scales, locs = hier_priors()
pairs = table.pairs()
by_ne = {}
for ip, (dg, ch) in enumerate(pairs):
    by_ne.setdefault(len(table.experiments(dg, ch)), []).append(ip)
K = 1000
n_all = 256 * len(pairs)
hier = {}
for ne, idxs in sorted(by_ne.items()):
    hp = HierPack([table.experiments(*pairs[i]) for i in idxs])
    hid = np.repeat(np.arange(len(idxs), dtype=np.int32), 256)
    th0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(hid), 1))
    hs = HierarchicalSampler(hp, hid, th0, pr, seed=ne, thinning=5, co_resident_chains=n_all - len(hid))
    hier[ne] = (hs, torch.empty((hs.n, K // 5, hs.d + 1), dtype=torch.float64, device="cuda"))


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def run(order, prio):
    streams = {ne: torch.cuda.Stream(priority=prio.get(ne, 0)) for ne in order}

    def go():
        ev = torch.cuda.Event(); ev.record()
        for ne in order:
            hs, hb = hier[ne]
            st = streams[ne]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                hs.run(K, samples=hb)
            done = torch.cuda.Event(); done.record(st)
            torch.cuda.current_stream().wait_event(done)
    return timed(go)


for order in [(3, 4, 5, 6), (6, 5, 4, 3), (4, 5, 6, 3), (3, 6, 5, 4), (4, 3, 5, 6), (5, 6, 3, 4)]:
    for prio_name, prio in [("equal", {}), ("Ne=3 high", {3: -1}), ("others high", {4: -1, 5: -1, 6: -1})]:
        t = run(order, prio)
        print("order %s  priority %-12s  %.2f ms -> %.3e chain-it/s" % (order, prio_name, t, n_all * K / (t * 1e-3)), flush=True)
