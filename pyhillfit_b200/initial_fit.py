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


# ---------------------------------------------------------------------------------------------
# The same fit for many datasets at once.  A command line that fits all 210 Crumb pairs spent more wall time in
# 210 scalar Nelder-Mead runs (and 740 per-experiment ones for the hierarchical start) than in the GPU sampler; here
# every numpy operation works on all datasets together: one vectorised grid, then Nelder-Mead with scipy's
# coefficients, initial simplex, acceptance rules and stopping test applied to every dataset under a mask.
# ---------------------------------------------------------------------------------------------
def _pad(datasets):
    n, width = len(datasets), max(len(c) for c, _ in datasets)
    C, Y, W = np.ones((n, width)), np.zeros((n, width)), np.zeros((n, width))
    for k, (c, y) in enumerate(datasets):
        C[k, :len(c)], Y[k, :len(c)], W[k, :len(c)] = c, y, 1.0
    return C, Y, W


def _ss_batch(pic50, hill, C, Y, W):
    """sum_of_square_diffs for every dataset: pic50, hill broadcast against the leading axes of C."""
    with np.errstate(all="ignore"):
        curve = 100. * (1. - 1. / (1. + (C / 10 ** (6 - pic50[..., None])) ** hill[..., None]))
        return np.sum(W * (curve - Y) ** 2, axis=-1)


def _nelder_mead_batch(f, x0, xatol, fatol, maxiter):
    """scipy.optimize.minimize(method='Nelder-Mead') (non-adaptive) for a batch: x0 [n, d] -> (x [n, d], f [n]).
    f(x, idx) evaluates the objective of problems idx at the points x [len(idx), d]; converged problems leave the
    working set, so a few slow ones (flat directions never meet xatol and run to maxiter) do not cost n each."""
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    n, d = x0.shape
    everyone = np.arange(n)
    sim = np.repeat(x0[:, None, :], d + 1, axis=1)                  # [n, d+1, d]
    for k in range(d):
        y = x0[:, k]
        sim[:, k + 1, k] = np.where(y != 0, 1.05 * y, 0.00025)
    fsim = np.stack([f(sim[:, j], everyone) for j in range(d + 1)], axis=1)   # [n, d+1]
    order = np.argsort(fsim, axis=1, kind="stable")
    sim, fsim = sim[everyone[:, None], order], fsim[everyone[:, None], order]
    out_x, out_f = sim[:, 0].copy(), fsim[:, 0].copy()
    idx = everyone
    for _ in range(maxiter):
        done = (np.max(np.abs(sim[:, 1:] - sim[:, :1]), axis=(1, 2)) <= xatol) & \
               (np.max(np.abs(fsim[:, :1] - fsim[:, 1:]), axis=1) <= fatol)
        if done.any():
            out_x[idx[done]], out_f[idx[done]] = sim[done, 0], fsim[done, 0]
            keep = ~done
            sim, fsim, idx = sim[keep], fsim[keep], idx[keep]
        if len(idx) == 0:
            break
        xbar = sim[:, :-1].sum(axis=1) / d
        worst = sim[:, -1]
        xr = (1 + rho) * xbar - rho * worst
        xe = (1 + rho * chi) * xbar - rho * chi * worst
        xoc = (1 + psi * rho) * xbar - psi * rho * worst
        xic = (1 - psi) * xbar + psi * worst
        m = len(idx)
        fall = f(np.concatenate([xr, xe, xoc, xic]), np.tile(idx, 4))        # one evaluation for the four candidates
        fxr, fxe, fxoc, fxic = fall[:m], fall[m:2 * m], fall[2 * m:3 * m], fall[3 * m:]
        f0, f2, fw = fsim[:, 0], fsim[:, -2], fsim[:, -1]
        expand = fxr < f0
        take_e = expand & (fxe < fxr)
        take_r = (expand & ~take_e) | (~expand & (fxr < f2))
        contract = ~expand & ~(fxr < f2)
        outside = contract & (fxr < fw)
        take_oc = outside & (fxoc <= fxr)
        inside = contract & ~(fxr < fw)
        take_ic = inside & (fxic < fw)
        shrink = (outside & ~take_oc) | (inside & ~take_ic)
        newx = np.where(take_e[:, None], xe, np.where(take_r[:, None], xr, np.where(take_oc[:, None], xoc, xic)))
        newf = np.where(take_e, fxe, np.where(take_r, fxr, np.where(take_oc, fxoc, fxic)))
        replace = ~shrink
        sim[replace, -1] = newx[replace]
        fsim[replace, -1] = newf[replace]
        if shrink.any():
            sh = np.nonzero(shrink)[0]
            sim[sh, 1:] = sim[sh, :1] + sigma * (sim[sh, 1:] - sim[sh, :1])
            for j in range(1, d + 1):
                fsim[sh, j] = f(sim[sh, j], idx[sh])
        order = np.argsort(fsim, axis=1, kind="stable")
        rows = np.arange(len(idx))[:, None]
        sim, fsim = sim[rows, order], fsim[rows, order]
    if len(idx):
        out_x[idx], out_f[idx] = sim[:, 0], fsim[:, 0]
    return out_x, out_f


def best_fit_batch(model, datasets, pic50_lower=PIC50_LOWER):
    """best_fit for a list of (concs, responses): -> theta0 [n, d], sum of squares [n].  Same objective, grid,
    re-parameterisation, tolerances and fall-backs as best_fit (tests/test_host_logic.py compares the two)."""
    if len(datasets) == 0:
        return np.zeros((0, 2 if model == 1 else 3)), np.zeros(0)
    datasets = [(np.asarray(c, dtype=float), np.asarray(y, dtype=float)) for c, y in datasets]
    C, Y, W = _pad(datasets)
    counts = W.sum(axis=1)
    PL = pic50_lower
    pic50_grid = np.linspace(PL, 12.0, 61)
    hill_grid = np.array([1.0]) if model == 1 else np.array([0.25, 0.5, 0.75, 1.0, 1.5, 2.0, 3.0, 5.0])
    H, P = np.meshgrid(hill_grid, pic50_grid, indexing="ij")          # hill outer, pIC50 inner: best_fit's loop order
    grid_ss = _ss_batch(P.reshape(-1)[None, :], H.reshape(-1)[None, :], C[:, None, :], Y[:, None, :], W[:, None, :])
    grid_ss = np.where(np.isnan(grid_ss), np.inf, grid_ss)
    pick = np.argmin(grid_ss, axis=1)                                 # first minimum, as the strict '<' keeps
    g_ss = grid_ss[np.arange(len(datasets)), pick]
    p0, h0 = P.reshape(-1)[pick], H.reshape(-1)[pick]
    if model == 1:
        f = lambda x, i: _ss_batch(x[:, 0] ** 2 + PL, np.ones(len(x)), C[i], Y[i], W[i])
        x, ss = _nelder_mead_batch(f, np.sqrt(p0 - PL)[:, None], 1e-10, 1e-12, 4000)
        pic50, hill = x[:, 0] ** 2 + PL, np.ones(len(x))
    else:
        f = lambda x, i: _ss_batch(x[:, 0] ** 2 + PL, x[:, 1] ** 2 + HILL_LOWER, C[i], Y[i], W[i])
        x, ss = _nelder_mead_batch(f, np.stack([np.sqrt(p0 - PL), np.sqrt(h0 - HILL_LOWER)], axis=1), 1e-10, 1e-12, 8000)
        pic50, hill = x[:, 0] ** 2 + PL, x[:, 1] ** 2 + HILL_LOWER
    worse = ~(ss <= g_ss)
    pic50, hill, ss = np.where(worse, p0, pic50), np.where(worse, h0, hill), np.where(worse, g_ss, ss)
    sigma = np.maximum(np.sqrt(ss / counts), 2e-3)
    hill = np.minimum(hill, 10.0)
    theta = np.stack([pic50, sigma], axis=1) if model == 1 else np.stack([pic50, hill, sigma], axis=1)
    return theta, ss


def best_fit_batch_gpu(model, datasets, pic50_lower=PIC50_LOWER, device=None):
    """best_fit_batch on the device (phf_best_fit_batch, one thread per dataset): same objective, grid, simplex rules,
    tolerances and fall-backs; `datasets` is a list of (concs, responses) or a tuple (offsets [n+1], concs, responses)
    of flat arrays.  -> theta0 [n, d], sum of squares [n] as numpy arrays.  Results agree with the host version to
    the rounding of pow() (tests/test_gpu_fit.py); a million datasets take about a second."""
    from . import _lib
    torch = _lib.require_cuda()
    if isinstance(datasets, tuple) and len(datasets) == 3 and np.ndim(datasets[0]) == 1 and not isinstance(datasets[0], tuple):
        offsets, concs, resp = (np.ascontiguousarray(a) for a in datasets)
        offsets = offsets.astype(np.int64)
    else:
        if len(datasets) == 0:
            return np.zeros((0, 2 if model == 1 else 3)), np.zeros(0)
        lens = np.array([len(c) for c, _ in datasets], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        concs = np.concatenate([np.asarray(c, dtype=np.float64) for c, _ in datasets])
        resp = np.concatenate([np.asarray(y, dtype=np.float64) for _, y in datasets])
    n = len(offsets) - 1
    if n <= 0:
        return np.zeros((0, 2 if model == 1 else 3)), np.zeros(0)
    if offsets[0] != 0 or np.any(np.diff(offsets) <= 0) or offsets[-1] != len(concs) or len(concs) != len(resp):
        raise ValueError("best_fit_batch_gpu: offsets must start at 0, increase strictly and end at len(concs)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        d_off = torch.from_numpy(offsets).to(dev)
        d_c = torch.from_numpy(np.asarray(concs, dtype=np.float64)).to(dev)
        d_y = torch.from_numpy(np.asarray(resp, dtype=np.float64)).to(dev)
        theta = torch.empty((n, 2 if model == 1 else 3), dtype=torch.float64, device=dev)
        ss = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.check(_lib.load().phf_best_fit_batch(model, n, _lib.ptr(d_off), _lib.ptr(d_c), _lib.ptr(d_y),
                                                  float(pic50_lower), _lib.ptr(theta), _lib.ptr(ss),
                                                  _lib.current_stream_ptr()), "phf_best_fit_batch")
        return theta.cpu().numpy(), ss.cpu().numpy()
