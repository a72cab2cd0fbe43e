"""Test helper: quantile comparison of pooled GPU chains with one reference chain, in units of the Monte-Carlo
standard error of a QUANTILE estimate (north_star: "posterior quantiles ... within Monte-Carlo standard error";
SURVEY.md section 8d: 3 x MCSE; the tests use 4 so that ~200 comparisons per suite keep a false-alarm rate < 2 %).

MCSE(q_p) = sqrt(p (1 - p) / ESS_p) / f(q_p), ESS_p = Geyer ESS of the indicator series 1[x_t <= q_p]
(pyhillfit_b200/ess.py) -- recorded for the reference chain by oracle/gen_golden_chains.py (`*_ess_q`), computed here
for the GPU chains (summed over chains); f(q_p) = density at the quantile, from the pooled GPU sample (central
difference of its quantile function over p +- 0.02).  Tolerance = n_sigma * sqrt(MCSE_ref^2 + MCSE_gpu^2).
"""
import numpy as np

from pyhillfit_b200.ess import ess_quantile_indicator, quantile_mcse

PS = np.array([0.05, 0.25, 0.50, 0.75, 0.95])


def quantile_z(samples, q_ref, ess_q_ref, ess_chains=16):
    """samples [n_chains, rows, d] (post-burn) -> z [5, d] = (q_gpu - q_ref) / sqrt(MCSE_ref^2 + MCSE_gpu^2)."""
    samples = np.asarray(samples, dtype=np.float64)
    nch, rows, d = samples.shape
    pooled = samples.reshape(-1, d)
    q = np.percentile(pooled, 100 * PS, axis=0)
    dp = 0.02
    dens = 2 * dp / np.maximum(np.percentile(pooled, 100 * (PS + dp), axis=0) -
                               np.percentile(pooled, 100 * (PS - dp), axis=0), 1e-300)
    use = min(ess_chains, nch)
    ess_gpu = np.array([[sum(ess_quantile_indicator(samples[c, :, j], q[i, j]) for c in range(use)) * nch / use
                         for j in range(d)] for i in range(len(PS))])
    p = PS[:, None]
    se = np.hypot(quantile_mcse(p, np.asarray(ess_q_ref, dtype=np.float64), dens), quantile_mcse(p, ess_gpu, dens))
    return (q - q_ref) / se, q, se


def assert_quantiles_within_mcse(samples, q_ref, ess_q_ref, n_sigma=4.0, what=""):
    z, q, se = quantile_z(samples, q_ref, ess_q_ref)
    assert np.all(np.abs(z) <= n_sigma), "%s: quantiles differ by up to %.2f MCSE (limit %.1f)\nz=%s\ngpu=%s\nref=%s" % (
        what, np.abs(z).max(), n_sigma, np.round(z, 2), q, q_ref)
    return z
