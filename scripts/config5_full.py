"""BASELINE config 5 at full size: 1 000 000 synthetic datasets (data/synthetic_data.csv shape) x 4 chains, model 2,
sharded over the ranks torchrun started (one process per GPU, no collective on the hot path).  Every rank generates and
packs its own shard (pyhillfit_b200.synthetic: dataset k is the same whatever the number of ranks), runs the fused
sampler with the thinned rows written to HBM, and the device time is the max over ranks.  Prints one JSON line.
Usage: torchrun ... scripts/config5_full.py [n_datasets] [iterations] [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from pyhillfit_b200 import dist as pd, synthetic
from pyhillfit_b200.packing import SinglePack
from pyhillfit_b200.sampler import SingleLevelSampler

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ws, rank, local = pd.world()
torch.cuda.set_device(local)
pd.init_process_group()
dist = torch.distributed
lo, hi = rank * n_total // ws, (rank + 1) * n_total // ws
t0 = time.time()
concs, Y, truth = synthetic.generate(hi - lo, offset=lo)
sp = SinglePack.from_uniform(concs, Y)
t_pack = time.time() - t0
ids = np.repeat(np.arange(sp.n_datasets, dtype=np.int32), 4)
s = SingleLevelSampler(2, sp, ids, 1.0, np.tile([6.0, 1.0, 6.0], (len(ids), 1)), variant="fit", seed=9,
                       chain_id_base=4 * lo, thinning=5)
rows = K // 5
buf = torch.empty((s.n, rows, 4), dtype=torch.float64, device="cuda")
for _ in range(3):                        # warm-up: 3 x K iterations (also burn-in: adaptation starts at t = 3000)
    s.run(K, samples=buf)
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    s.run(K, samples=buf)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
# size-independent property: the pooled posterior mean of pIC50 over a dataset's 4 chains recovers the generating
# value (20 points per dataset: the posterior sd is a few tenths of a log unit)
post = buf[:, :, 0].mean(dim=1).reshape(-1, 4).mean(dim=1).cpu().numpy()
err = np.abs(post - truth[:, 0])
stats = torch.tensor([np.median(err), float((err < 1.0).mean()), float(np.isfinite(buf[:, -1, 3].cpu().numpy()).mean())],
                     dtype=torch.float64, device="cuda")
if ws > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats /= ws
if rank == 0:
    sec = float(ms.item()) * 1e-3
    chains = 4 * n_total
    print(json.dumps({"config": "BASELINE config 5: %d synthetic datasets x 4 chains, model 2" % n_total, "n_gpus": ws,
                      "chains": chains, "chains_per_gpu": s.n, "iters_per_step": K, "steps": steps,
                      "value": chains * K * steps / sec, "unit": "chain-iterations/s",
                      "ms_per_step": sec * 1e3 / steps, "lanes": s.lanes, "block_threads": s.block_threads,
                      "scaling": "strong", "timing": "CUDA events, barrier before, max over ranks; samples written to HBM (%.1f GB per step per GPU)" % (buf.numel() * 8 / 1e9),
                      "generate_and_pack_s_per_rank": round(t_pack, 2),
                      "median_abs_error_posterior_mean_pIC50": float(stats[0]), "frac_within_1_log_unit": float(stats[1]),
                      "frac_finite_log_target": float(stats[2])}))
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
