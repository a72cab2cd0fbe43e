"""Effective sample size (the reference has no ESS estimator; SURVEY.md section 8d defines this one).

Geyer's initial-positive-sequence estimator on the FFT autocovariance of one scalar series:
ESS = n / (-1 + 2 * sum_k Gamma_k), Gamma_k = rho_{2k} + rho_{2k+1}, truncated at the first non-positive
Gamma_k.  `ess_min` is the minimum over parameter columns -- the figure ESS/s is quoted on.
The same function is applied to GPU chains and to CPU-reference chains.
"""
import numpy as np


def autocorr_fft(x):
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    x = x - x.mean()
    m = 1 << (2 * n - 1).bit_length()
    f = np.fft.rfft(x, m)
    acov = np.fft.irfft(f * np.conj(f), m)[:n] / n
    if acov[0] <= 0:
        return np.ones(1)
    return acov / acov[0]


def ess_geyer(x):
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if n < 4 or np.ptp(x) == 0:
        return float(n)
    rho = autocorr_fft(x)
    k = (len(rho) // 2) * 2
    pair = rho[0:k:2] + rho[1:k:2]
    neg = np.nonzero(pair <= 0)[0]
    stop = neg[0] if len(neg) else len(pair)
    tau = -1.0 + 2.0 * pair[:stop].sum()
    tau = max(tau, 1.0 / n)
    return float(min(n / tau, n * 1.0))


def ess_min(chain):
    """min over columns of a [rows, d] array."""
    chain = np.asarray(chain, dtype=np.float64)
    return min(ess_geyer(chain[:, j]) for j in range(chain.shape[1]))


def mcse_mean(x):
    x = np.asarray(x, dtype=np.float64)
    return float(x.std(ddof=1) / np.sqrt(ess_geyer(x)))


def ess_quantile_indicator(x, q):
    """ESS of the indicator series 1[x_t <= q]: Var(F_hat(q)) = p (1 - p) / ESS, p = F(q) -- the effective sample size
    that governs the Monte-Carlo error of a quantile estimate (a tail indicator mixes more slowly than the mean)."""
    return ess_geyer((np.asarray(x, dtype=np.float64) <= q).astype(np.float64))


def quantile_mcse(p, ess_p, density):
    """Monte-Carlo standard error of the p-quantile estimate: sqrt(p (1 - p) / ESS_p) / f(q_p)."""
    return np.sqrt(p * (1.0 - p) / ess_p) / density
