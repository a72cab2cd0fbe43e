"""Developer probe: does a shard's slowest dataset set the time of a latency-bound (sharded) TI sweep?  Runs ti.run_ti on
26 Crumb pairs (one GPU's share of eight) with and without five-dose pairs, and with the five-dose pairs given 4 lanes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200 import ti
table = Table("crumb_data")
pairs = table.pairs()
nd = {p: len(np.unique(table.concat(*p)[0])) for p in pairs}
ncens = {p: int(((table.concat(*p)[1] == 0) | (table.concat(*p)[1] == 100)).sum()) for p in pairs}
four = [p for p in pairs if nd[p] == 4]
five = [p for p in pairs if nd[p] == 5]
light = sorted(four, key=lambda p: ncens[p])[:26]
heavy = sorted(four, key=lambda p: -ncens[p])[:26]
sets = {"26 four-dose pairs, fewest censored responses": light, "26 four-dose pairs, most censored responses": heavy,
        "24 light four-dose pairs + 2 five-dose pairs": light[:24] + five[:2], "mixed: 13 light + 13 heavy": light[:13] + heavy[:13]}
for name, ps in sets.items():
    data = [table.concat(*p) for p in ps]
    ti.run_ti(data, iterations=20000, segment=20000)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ti.run_ti(data, iterations=100000, segment=100000)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-52s %d chains  %.3f s per 1e5 iterations  (%.0f cycles per iteration)  lanes %s speculation %s" % (
        name, out["chains_local"], dt, dt * 1.965e9 / 1e5, out["lanes"], out["speculation"]), flush=True)
