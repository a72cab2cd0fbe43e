"""The reference's full thermodynamic-integration workload in one go: 210 Crumb pairs x models {1,2} x the 41-point
ladder at the reference defaults (500 000 iterations, thinning 5, burn-in 1/4) -> ln p(y|M1), ln p(y|M2), B12 per
pair.  The reference needs 210 x 2 PyHillTemp runs + 210 compute_bayes_factors runs (extrapolated 82 h on 8 cores,
SURVEY.md section 6).  Usage: full_ti_sweep.py [iterations] [replicates]   (under torchrun for several GPUs)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from _data import Table
from pyhillfit_b200 import dist as pd, ti
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ws, rank, local = pd.world()
torch.cuda.set_device(local)
pd.init_process_group()
table = Table("crumb_data")
pairs = table.pairs()
data = [table.concat(*p) for p in pairs]
if ws > 1:     # bring the NCCL communicator up before the clock starts (its first collective takes about a second)
    torch.distributed.all_reduce(torch.zeros(1, device="cuda"))
torch.cuda.synchronize()
t0 = time.time()
out = ti.run_ti(data, replicates=reps, iterations=iters, thinning=5, burn_in_fraction=4, seed=1, segment=100000)
torch.cuda.synchronize()
dt = time.time() - t0
ph = torch.tensor([out.get("sample_seconds", 0.0), out["gather_seconds"]], dtype=torch.float64, device="cuda")
if ws > 1:
    torch.distributed.all_reduce(ph, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
    n = len(pairs) * 2 * 41 * reps
    print("%d chains x %d iterations on %d GPU(s): %.2f s wall (%.3e chain-iterations/s incl. packing and launch overheads)"
          % (n, iters, ws, dt, n * iters / dt))
    print("  sampling phase %.3f s (max over ranks, %.3e chain-iterations/s), all-gather + trapezium rule %.4f s"
          % (float(ph[0]), n * iters / float(ph[0]), float(ph[1])))
    b = out["B12"]
    print("B12: min %.3g median %.3g max %.3g; pairs favouring model 2 (B12 < 1): %d of %d" % (b.min(), np.median(b), b.max(), int((b < 1).sum()), len(b)))
    for i in (0, 1, 2):
        print("  %s/%s  ln p(y|M1) %.4f  ln p(y|M2) %.4f  B12 %.4g" % (pairs[i][0], pairs[i][1], out["log_py"][1][i], out["log_py"][2][i], b[i]))
if ws > 1:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
