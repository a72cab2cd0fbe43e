"""Developer sweep: (lanes per evaluation E) x (speculation depth S) x CTA size for small launches -- the latency
regime of a sharded thermodynamic-integration sweep.  Both models run co-resident on two streams (as ti.run_ti does);
prints ms per 20 000 iterations and chain-iterations/s, and checks that every combination with the same E gives the
same bits.  Usage: python scripts/spec_sweep.py [chains_per_model ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200 import ti
from pyhillfit_b200.packing import SinglePack
from pyhillfit_b200.sampler import SingleLevelSampler

table = Table("crumb_data")
pack = SinglePack([table.concat(d, c) for d, c in table.pairs()])
temps = ti.temperature_ladder(40, 3)
ids_all, tt_all = ti.build_chain_list(210, temps, 1)           # 8610 chains per model
K = 20000
sizes = [int(a) for a in sys.argv[1:]] or [1076, 2152, 4305, 8610]
for n in sizes:
    if n <= len(ids_all):
        lo = (len(ids_all) - n) // 2
        ids, tt = ids_all[lo:lo + n], tt_all[lo:lo + n]
    else:
        reps = -(-n // len(ids_all))
        ids, tt = np.sort(np.tile(ids_all, reps)[:n]), np.tile(tt_all, reps)[:n]
    ref = {}
    print("== %d chains per model (x 2 models co-resident)" % n)
    for E, S, bt in [(1, 1, 0), (2, 1, 0), (4, 1, 0), (1, 2, 0), (1, 4, 0), (2, 2, 0), (2, 4, 0), (2, 8, 0), (4, 2, 0), (4, 4, 0),
                     (4, 8, 0), (4, 4, 32), (4, 4, 128), (2, 4, 32), (2, 4, 128), (4, 8, 32), (2, 8, 32)]:
        ss = {}
        for m in (1, 2):
            d = 2 if m == 1 else 3
            ss[m] = SingleLevelSampler(m, pack, ids, tt, np.ones((n, d)), variant="temp", seed=1, thinning=5,
                                       burn_rows=1000, lanes=E, speculation=S, block_threads=bt, chain_id_base=m << 40)
        st = {m: torch.cuda.Stream() for m in (1, 2)}

        def both():
            cur = torch.cuda.current_stream()
            for m in (1, 2):
                st[m].wait_stream(cur)
                with torch.cuda.stream(st[m]):
                    ss[m].run(K, keep=False)
            for m in (1, 2):
                cur.wait_stream(st[m])
        both()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); both(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        state = np.concatenate([ss[m].state.cpu().numpy().ravel() for m in (1, 2)])
        same = ""
        if S == 1:
            ref[E] = state
        elif E in ref:
            same = "bit-identical to S=1" if np.array_equal(ref[E], state) else "DIFFERS from S=1 (max rel %.2e)" % np.max(
                np.abs(ref[E] - state) / np.maximum(np.abs(ref[E]), 1e-300))
        print("E=%d S=%d block=%3d: %8.2f ms / %d its  %.3e chain-it/s  %6.0f cycles/it  acc %.3f  %s" % (
            E, S, ss[2].block_threads or bt, ms, K, 2 * n * K / (ms * 1e-3), ms * 1e-3 * 1.965e9 / K,
            float(ss[2].acceptance().mean()), same))
        del ss
