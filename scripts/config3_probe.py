"""Developer probe: BASELINE config 3 (hierarchical, 210 pairs x 256 chains, four launches by number of experiments)
under different kernel choices per launch.  usage: config3_probe.py  (prints ms per 1000 iterations, alone and all
four launches concurrent on four streams)"""
import os, sys
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


def build(choice):
    """choice: {ne: (lanes, hint)}; lanes 1 = thread kernel (hint 1: covariance in shared memory, 0: in L2), 16/32 lane kernel"""
    out = []
    for ne, idxs in sorted(by_ne.items()):
        hp = HierPack([table.experiments(*pairs[i]) for i in idxs])
        hid = np.repeat(np.arange(len(idxs), dtype=np.int32), 256)
        th0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(hid), 1))
        lanes, hint = choice[ne][:2]
        hs = HierarchicalSampler(hp, hid, th0, pr, seed=ne, thinning=5, lanes=lanes,
                                 block_threads=choice[ne][2] if len(choice[ne]) > 2 else 0)
        hs.occupancy_hint = hint
        hb = torch.empty((hs.n, K // 5, hs.d + 1), dtype=torch.float64, device="cuda")
        out.append((hs, hb, torch.cuda.Stream()))
    return out


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


choices = {
    "round-1 default (Ne=3 thread/smem, others lane)": {3: (1, 1), 4: (16, 0), 5: (16, 0), 6: (32, 0)},
    "Ne=3,4 thread, 5,6 lane": {3: (1, 1), 4: (1, 1), 5: (16, 0), 6: (32, 0)},
    "Ne=3 thread, 4 quad, 5,6 lane": {3: (1, 1), 4: (4, 0), 5: (16, 0), 6: (32, 0)},
    "Ne=3 thread, 4,5 quad, 6 lane": {3: (1, 1), 4: (4, 0), 5: (4, 0), 6: (32, 0)},
    "Ne=3,4,5 quad, 6 lane": {3: (4, 0), 4: (4, 0), 5: (4, 0), 6: (32, 0)},
    "library's choice": {3: (0, 0), 4: (0, 0), 5: (0, 0), 6: (0, 0)},
    "Ne=3,4 thread, 5,6 lane in 32-thread CTAs": {3: (1, 1), 4: (1, 1), 5: (16, 0, 32), 6: (32, 0, 32)},
    "Ne=3,4 thread, 5,6 lane in 64-thread CTAs": {3: (1, 1), 4: (1, 1), 5: (16, 0, 64), 6: (32, 0, 64)},
    "Ne=3 thread, 4,5,6 lane in 32-thread CTAs": {3: (1, 1), 4: (16, 0, 32), 5: (16, 0, 32), 6: (32, 0, 32)},
    "all thread": {3: (1, 1), 4: (1, 1), 5: (1, 1), 6: (1, 1)},
}
only = sys.argv[1:]
for name, ch in choices.items():
    if only and not any(o in name for o in only):
        continue
    hier = build(ch)
    alone = [timed(lambda hs=hs, hb=hb: hs.run(K, samples=hb)) for hs, hb, _ in hier]

    def all_at_once():
        ev = torch.cuda.Event(); ev.record()
        for hs, hb, st in hier:
            st.wait_event(ev)
            with torch.cuda.stream(st):
                hs.run(K, samples=hb)
            done = torch.cuda.Event(); done.record(st)
            torch.cuda.current_stream().wait_event(done)
    tot = timed(all_at_once)
    n = sum(h[0].n for h in hier)
    print("%-48s alone ms %s  sum %.1f  concurrent %.1f ms -> %.3e chain-it/s" % (
        name, " ".join("%.1f" % a for a in alone), sum(alone), tot, n * K / (tot * 1e-3)), flush=True)
    del hier
