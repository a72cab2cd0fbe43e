"""Least-squares start point for the single-level sampler (host side, not on the hot path).

The reference finds theta0 with CMA-ES on the sum of squared residuals over (pIC50, Hill) in an x**2
re-parameterisation, then sets sigma0 = sqrt(SS/N) (python/PyHillFit.py:93-102, 699-735).  `cma` is a
third-party package that is neither vendored by the reference nor installed here, so this module minimises
the SAME objective with a deterministic coarse grid + Nelder-Mead polish ("parity unpinned": only the
start point and best_fit_params.txt depend on it; the posterior does not).
"""
import numpy as np
from scipy.optimize import minimize

PIC50_LOWER = -3.0   # dr.pic50_exp_lower
HILL_LOWER = 0.0     # dr.hill_uniform_lower


def _curve(concs, hill, pic50):
    with np.errstate(all="ignore"):
        return 100. * (1. - 1. / (1. + (concs / 10 ** (6 - pic50)) ** hill))


def sum_of_square_diffs(params, concs, responses):
    """python/PyHillFit.py:93-97"""
    pic50, hill = params
    return float(np.sum((_curve(concs, hill, pic50) - responses) ** 2))


def best_fit(model, concs, responses, pic50_lower=PIC50_LOWER):
    """-> theta0: (pIC50, sigma) for model 1, (pIC50, Hill, sigma) for model 2, and the sum of squares."""
    concs = np.asarray(concs, dtype=float)
    responses = np.asarray(responses, dtype=float)
    PL = pic50_lower
    pic50_grid = np.linspace(PL, 12.0, 61)
    hill_grid = np.array([1.0]) if model == 1 else np.array([0.25, 0.5, 0.75, 1.0, 1.5, 2.0, 3.0, 5.0])
    best = (np.inf, None)
    for h in hill_grid:
        for p in pic50_grid:
            ss = sum_of_square_diffs((p, h), concs, responses)
            if ss < best[0]:
                best = (ss, (p, h))
    p0, h0 = best[1]
    if model == 1:
        obj = lambda x: sum_of_square_diffs((x[0] ** 2 + PL, 1.0), concs, responses)
        res = minimize(obj, [np.sqrt(p0 - PL)], method="Nelder-Mead",
                       options=dict(xatol=1e-10, fatol=1e-12, maxiter=4000))
        pic50, hill, ss = res.x[0] ** 2 + PL, 1.0, res.fun
    else:
        obj = lambda x: sum_of_square_diffs((x[0] ** 2 + PL, x[1] ** 2 + HILL_LOWER), concs, responses)
        res = minimize(obj, [np.sqrt(p0 - PL), np.sqrt(h0 - HILL_LOWER)], method="Nelder-Mead",
                       options=dict(xatol=1e-10, fatol=1e-12, maxiter=8000))
        pic50, hill, ss = res.x[0] ** 2 + PL, res.x[1] ** 2 + HILL_LOWER, res.fun
    if ss > best[0]:
        pic50, hill, ss = p0, h0, best[0]
    sigma = np.sqrt(ss / len(responses))  # initial_sigma, python/PyHillFit.py:101-102
    sigma = max(sigma, 2e-3)              # a perfect fit would start at the prior's edge
    hill = min(hill, 10.0)
    theta = np.array([pic50, sigma]) if model == 1 else np.array([pic50, hill, sigma])
    return theta, float(ss)
