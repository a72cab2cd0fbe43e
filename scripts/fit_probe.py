"""Time phf_best_fit_batch on BASELINE config 5's synthetic datasets (kernel only, CUDA events) and the host
best_fit_batch on a sample of them.  Usage: python scripts/fit_probe.py [n_datasets] [host_sample]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyhillfit_b200 import _lib, synthetic
from pyhillfit_b200.initial_fit import best_fit_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
host_n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
c20, Y, truth = synthetic.generate(n)
L = _lib.load()
dev = torch.device("cuda", 0)
off = torch.arange(n + 1, dtype=torch.int64, device=dev) * 20
dc = torch.from_numpy(np.tile(c20, n)).to(dev)
dy = torch.from_numpy(Y.reshape(-1)).to(dev)
out = {}
for model in (1, 2):
    th = torch.empty((n, model + 1), dtype=torch.float64, device=dev)
    ss = torch.empty(n, dtype=torch.float64, device=dev)
    ms = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.phf_best_fit_batch(model, n, off.data_ptr(), dc.data_ptr(), dy.data_ptr(), -3.0, th.data_ptr(),
                                        ss.data_ptr(), torch.cuda.current_stream().cuda_stream), "fit")
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = time.time()
    hth, hss = best_fit_batch(model, [(c20, Y[k]) for k in range(host_n)])
    host_s = time.time() - t
    g = ss[:host_n].cpu().numpy()
    out["model_%d" % model] = {"n_datasets": n, "gpu_ms": ms, "datasets_per_s_gpu": n / (min(ms) * 1e-3),
                               "host_sample": host_n, "host_s": host_s, "datasets_per_s_host": host_n / host_s,
                               "max_rel_ss_diff_vs_host": float(np.max(np.abs(g - hss) / np.maximum(hss, 1.0)))}
print(json.dumps(out, indent=1))
