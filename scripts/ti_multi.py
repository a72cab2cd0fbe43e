"""Thermodynamic-integration sweep over a few Crumb pairs on however many ranks torchrun started (developer /
multi-GPU check): prints log p(y|M) and B12 per pair with full precision, so a 1-rank and an N-rank run can be
diffed -- with a fixed lane count they must agree bit for bit (chains keep their global Philox ids)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200 import dist as pd, ti
ws, rank, local = pd.world()
torch.cuda.set_device(local)
pd.init_process_group()
table = Table("crumb_data")
pairs = table.pairs()[:int(sys.argv[1]) if len(sys.argv) > 1 else 12]
out = ti.run_ti([table.concat(*p) for p in pairs], replicates=2, iterations=20000, thinning=5, seed=3, lanes=1)
if rank == 0:
    for i, p in enumerate(pairs):
        print("%s/%s %r %r %r" % (p[0], p[1], float(out["log_py"][1][i]), float(out["log_py"][2][i]), float(out["B12"][i])))
if ws > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
