"""Generate tests/golden/cdf_golden.npz by executing the UNMODIFIED reference function
construct_posterior_predictive_cdfs (python/construct_hierarchical_cdfs.py:32-58) via oracle/ref_shim.py on a
fixed set of (alpha, beta, mu, s) rows.  Build container only (needs /root/reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def rows(n=3000, seed=2016):
    rng = np.random.default_rng(seed)
    r = np.stack([rng.uniform(0.3, 2.0, n), rng.uniform(2.05, 10.0, n), rng.uniform(3.0, 9.0, n),
                  np.exp(rng.uniform(np.log(0.012), np.log(1.5), n))], 1)
    r[0] = [1.0, 2.5, 6.0, 0.011]      # very narrow logistic: exp overflow side of the grid
    r[1] = [0.05, 40.0, -1.5, 3.0]     # very steep log-logistic
    return r


if __name__ == "__main__":
    f = ref_shim.load_construct_cdfs()
    r = rows()
    with np.errstate(all="ignore"):
        hx, hc, px, pc, hp, pp = f(r[:, 0], r[:, 1], r[:, 2], r[:, 3])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cdf_golden.npz"), rows=r, hill_x=hx, hill_cdf=hc,
                        pic50_x=px, pic50_cdf=pc, hill_pdf=hp, pic50_pdf=pp)
    print("cdf_golden.npz", r.shape, hc[[0, 100, 500]], pc[[0, 250, 500]])
