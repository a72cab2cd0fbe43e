"""CPU oracle: numpy/scipy restatement of PyHillFit's MCMC hot path.

TEST INFRASTRUCTURE ONLY -- not part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file; ``pyhillfit_b200/`` never does (the product path fails
loudly when the CUDA library is missing).

Every function restates one reference function and cites it (paths relative to
``/root/reference``).  It deliberately issues the *same* numpy / scipy calls in the
same order as the reference so that (a) values agree with the reference bit for
bit on the same numpy/scipy, and (b) its run time is a faithful stand-in for the
reference's CPU cost (``bench.py --impl reference``; the reference itself is
Python 2 and cannot travel to the GPU box).

Pinning: ``tests/golden/*.npz`` hold outputs of the UNMODIFIED reference executed in
the build container through ``oracle/ref_shim.py`` (generator:
``oracle/gen_golden.py``).  ``tests/test_oracle_golden.py`` checks this file and the C
restatement (``oracle/hill_oracle.c``) against them.  The reference has no tests of
its own for this path.

Differences from the reference, all explicit:
  * the model is an argument, not module-global state (doseresponse.py:250-296);
  * the AM loops can draw from either ``numpy.random.RandomState`` exactly as the
    reference does (``rng="numpy"``) or from the counter-based Philox stream the CUDA
    kernels use (``rng="philox"``; Cholesky factor instead of numpy's SVD factor) so
    that GPU trajectories can be followed step by step.
"""
import math

import numpy as np
import scipy.stats as st

# ----------------------------------------------------------------------------
# constants -- python/doseresponse.py:12-28
# ----------------------------------------------------------------------------
sigma_uniform_lower = 1e-3
pic50_exp_rate = 0.2
pic50_exp_lower = -3.
hill_uniform_lower = 0.
hill_uniform_upper = 10.
sigma_shape = 5.
sigma_mode = 6.
sigma_loc = 1e-3
sigma_scale = (sigma_mode - sigma_loc) / (sigma_shape - 1.)
n = 40
c = 3

NUM_PARAMS = {1: 2, 2: 3}  # doseresponse.py:259,275


# ----------------------------------------------------------------------------
# Hill curve -- python/doseresponse.py:84-88
# ----------------------------------------------------------------------------
def dose_response_model(dose, hill, IC50):
    return 100. * (1. - 1. / (1. + (1. * dose / IC50) ** hill))


def pic50_to_ic50(pic50):
    return 10 ** (6 - pic50)


# ----------------------------------------------------------------------------
# priors -- python/doseresponse.py:151-156, 166-184, 304-317
# ----------------------------------------------------------------------------
def log_pic50_exponential(x):
    if x < pic50_exp_lower:
        return -np.inf
    return -pic50_exp_rate * x


def log_gamma_prior(x, shape_param, scale_param, loc_params):
    if np.any(x < loc_params):
        return -np.inf
    return (shape_param - 1) * np.log(x - loc_params) - (x - loc_params) / scale_param


def log_priors(model, params):
    if model == 1:
        pic50, sigma = params
    else:
        pic50, hill, sigma = params
        if (hill < hill_uniform_lower) or (hill > hill_uniform_upper):
            return -np.inf
    return log_pic50_exponential(pic50) + log_gamma_prior(sigma, sigma_shape, sigma_scale, sigma_loc)


# ----------------------------------------------------------------------------
# censored likelihoods -- python/doseresponse.py:203-248, 299-301, 187-189
# ----------------------------------------------------------------------------
def compute_pi_bit_of_log_likelihood(y):
    return 0.5 * len(y) * np.log(2 * np.pi)


def log_data_likelihood(model, y, where_y_0, where_y_100, where_y_other, concs, params, t, pi_bit):
    if t == 0:
        return 0
    if model == 1:
        pic50, sigma = params
        hill = 1
    else:
        pic50, hill, sigma = params
    if sigma <= sigma_uniform_lower:
        return -np.inf
    predicted = dose_response_model(concs, hill, pic50_to_ic50(pic50))
    y_0_sum = np.sum(st.norm.logcdf(0, predicted[where_y_0], sigma))
    y_100_sum = np.sum(st.norm.logsf(100, predicted[where_y_100], sigma))
    temp_1 = where_y_other.sum() * np.log(sigma)
    temp_2 = np.sum((y[where_y_other] - predicted[where_y_other]) ** 2 / (2. * sigma ** 2))
    return t * (y_0_sum + y_100_sum - pi_bit - temp_1 - temp_2)


def log_target(model, y, where_y_0, where_y_100, where_y_other, concs, params, t, pi_bit):
    return (log_data_likelihood(model, y, where_y_0, where_y_100, where_y_other, concs, params, t, pi_bit)
            + log_priors(model, params))


def masks(responses):
    """PyHillFit.py:675-677 / PyHillTemp.py:138-140 (a response outside [0,100] is in no mask)."""
    return responses == 0, responses == 100, (0 < responses) & (responses < 100)


def concat_experiments(experiments):
    """PyHillFit.py:661-665: experiments concatenated in list order."""
    concs = np.concatenate([np.asarray(e)[:, 0] for e in experiments])
    responses = np.concatenate([np.asarray(e)[:, 1] for e in experiments])
    return concs, responses


# ----------------------------------------------------------------------------
# hierarchical target -- python/PyHillFit.py:113-154, 173-193, 301, 340-364
# ----------------------------------------------------------------------------
def hier_log_data_likelihood(hill_is, pic50_is, sigma, experiments):
    answer = 0.
    for i in range(len(experiments)):
        ic50 = pic50_to_ic50(pic50_is[i])
        concs = experiments[i][:, 0]
        data = experiments[i][:, 1]
        model_responses = dose_response_model(concs, hill_is[i], ic50)
        exp_bit = np.sum((data - model_responses) ** 2) / (2 * sigma ** 2)
        truncated_scale = np.sum(np.log(st.norm.cdf(100, model_responses, sigma)
                                        - st.norm.cdf(0, model_responses, sigma)))
        answer -= (len(concs) * np.log(sigma) + exp_bit + truncated_scale)
    return answer


def log_hill_i_log_logistic_likelihood(x, alpha, beta):
    return np.log(beta) - beta * np.log(alpha) + (beta - 1.) * np.log(x) - 2 * np.log(1 + (x / alpha) ** beta)


def log_pic50_i_logistic_likelihood(x, mu, s):
    temp_bit = (x - mu) / s
    return -temp_bit - np.log(s) - 2 * np.log(1 + np.exp(-temp_bit))


HIER_PIC50_LOWER = -2.  # PyHillFit.py:215


def hier_prior_constants():
    """(shapes, scales, locs) of the Gamma hyper-priors -- PyHillFit.py:301, 340-364."""
    locs = np.array([0., 2., -4, 0.01, sigma_loc])
    elkins_hill_alphas = np.array([1.188, 1.744, 1.530, 0.930, 0.605, 1.325, 1.179, 0.979, 1.790, 1.708, 1.586,
                                   1.469, 1.429, 1.127, 1.011, 1.318, 1.063])
    elkins_hill_betas = 1. / np.array([0.0835, 0.1983, 0.2089, 0.1529, 0.1206, 0.2386, 0.2213, 0.2263, 0.1784,
                                       0.1544, 0.2486, 0.2031, 0.2025, 0.1510, 0.1837, 0.1677, 0.0862])
    elkins_pic50_mus = np.array([5.235, 5.765, 6.060, 5.315, 5.571, 7.378, 7.248, 5.249, 6.408, 5.625, 7.321,
                                 6.852, 6.169, 6.217, 5.927, 7.414, 4.860])
    elkins_pic50_sigmas = np.array([0.0760, 0.1388, 0.1459, 0.2044, 0.1597, 0.2216, 0.1856, 0.1560, 0.1034,
                                    0.1033, 0.1914, 0.1498, 0.1464, 0.1053, 0.1342, 0.1808, 0.0860])
    modes = np.array([np.mean(elkins_hill_alphas), np.mean(elkins_hill_betas) - 2., np.mean(elkins_pic50_mus),
                      np.mean(elkins_pic50_sigmas), sigma_mode])
    shapes = np.array([5., 2.5, 7.5, 2.5, sigma_shape])
    scales = (modes - locs) / (shapes - 1.)
    return shapes, scales, locs


def hier_log_target(experiments, theta, shapes, scales, locs):
    if np.any(theta[:4] <= locs[:4]):
        return -np.inf
    alpha, beta, mu, s = theta[:4]
    pic50_is = theta[4:-1:2]
    hill_is = theta[5:-1:2]
    sigma = theta[-1]
    if np.any(hill_is < 0) or np.any(pic50_is < HIER_PIC50_LOWER) or (sigma <= locs[-1]):
        return -np.inf
    total = hier_log_data_likelihood(hill_is, pic50_is, sigma, experiments)
    total += np.sum(log_hill_i_log_logistic_likelihood(hill_is, alpha, beta))
    total += np.sum(log_pic50_i_logistic_likelihood(pic50_is, mu, s))
    total += np.sum(log_gamma_prior(theta[[0, 1, 2, 3, -1]], shapes, scales, locs))
    return total


# ----------------------------------------------------------------------------
# thermodynamic integration tail -- doseresponse.py:27-28,192-193; PyHillTemp.py:151;
# compute_bayes_factors.py:11-27, 83-94
# ----------------------------------------------------------------------------
def temperature_ladder(n_=n, c_=c):
    return (np.arange(n_ + 1.) / n_) ** c_


def trapezium_rule(x, y):
    return 0.5 * np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]))


def compute_log_py_approxn(model, chain, y, w0, w100, wother, concs, pi_bit):
    """Mean over chain rows of the temperature-1 log-likelihood (compute_bayes_factors.py:11-27)."""
    d = NUM_PARAMS[model]
    total = 0.
    for it in range(chain.shape[0]):
        total += log_data_likelihood(model, y, w0, w100, wother, concs, chain[it, :d], 1, pi_bit)
    return total / chain.shape[0]


def bayes_factor_12(temps, log_p_ys_m1, log_p_ys_m2):
    return np.exp(trapezium_rule(temps, log_p_ys_m1) - trapezium_rule(temps, log_p_ys_m2))


# ----------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller: the stream contract shared with the CUDA kernels
# (replaces numpy's MT19937; SURVEY.md section 2 "Third-party: numpy.random").
#   call(seed, chain, t, j) -> 4 x uint32 with key=(seed_lo, seed_hi),
#   counter=(t, j, chain_lo, chain_hi).
#   j = 0 : words 0,1 -> accept uniform (53 bit);  words 2,3 -> normal pair 0 (z0, z1)
#   j >= 1: words 0,1 -> normal pair 2j-1;  words 2,3 -> normal pair 2j
#   normal pair from words (a, b): r = sqrt(-2 ln((a+1) 2^-32)), phi = pi * (b 2^-31);
#                                  z_even = r cos(phi), z_odd = r sin(phi)
# ----------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox_call(seed, chain, t, j):
    return philox4x32_10((t & _MASK, j & _MASK, chain & _MASK, (chain >> 32) & _MASK),
                         (seed & _MASK, (seed >> 32) & _MASK))


def uniform53(w0, w1):
    return (float(((w0 << 32) | w1) >> 11) + 0.5) * 2.0 ** -53


def box_muller(a, b):
    r = math.sqrt(-2.0 * math.log((a + 1.0) * 2.0 ** -32))
    phi = math.pi * (b * 2.0 ** -31)
    return r * math.cos(phi), r * math.sin(phi)


def philox_draw(seed, chain, t, d):
    """(u_accept, z[0:d]) for iteration t of one chain."""
    z = []
    w = philox_call(seed, chain, t, 0)
    u = uniform53(w[0], w[1])
    z.extend(box_muller(w[2], w[3]))
    j = 1
    while len(z) < d:
        w = philox_call(seed, chain, t, j)
        z.extend(box_muller(w[0], w[1]))
        if len(z) < d:
            z.extend(box_muller(w[2], w[3]))
        j += 1
    return u, np.array(z[:d])


# ----------------------------------------------------------------------------
# adaptive-Metropolis loops -- PyHillFit.py:748-751,787-856 (variant "fit"),
# PyHillFit.py:431-511 (variant "hier"), PyHillTemp.py:57-125 (variant "temp").
# SURVEY.md section 3.5 tabulates the differences.
# ----------------------------------------------------------------------------
PIVOT_FLOOR = 1e-12
# A zero variance in Sigma0 (a theta0 component of exactly 0 under 0.05 diag|theta0|) freezes that coordinate in the
# reference (numpy's SVD-based multivariate_normal draws it with zero spread).  The Cholesky-based Philox path stores
# such a diagonal entry as COV0_DIAG_FLOOR instead (include/pyhillfit_b200.h: PHF_COV0_DIAG_FLOOR): the coordinate
# moves by ~1e-30, i.e. stays where it is, and the factor stays finite.
COV0_DIAG_FLOOR = 1e-60


def guarded_cholesky(a):
    """Lower Cholesky factor with every pivot floored at PIVOT_FLOOR * a[j,j] (see oracle/hill_oracle.c: the
    adapted covariance can be numerically singular; numpy's SVD-based multivariate_normal tolerates that)."""
    a = np.asarray(a, dtype=float)
    d = a.shape[0]
    l = np.zeros((d, d))
    for j in range(d):
        s = a[j, j] - np.dot(l[j, :j], l[j, :j])
        fl = PIVOT_FLOOR * a[j, j]
        if not s > fl:
            s = fl
        if not s > 0:
            raise np.linalg.LinAlgError("non-positive diagonal")
        l[j, j] = math.sqrt(s)
        for i in range(j + 1, d):
            l[i, j] = (a[i, j] - np.dot(l[i, :j], l[j, :j])) / l[j, j]
    return l


def am_defaults(variant, theta0):
    """(cov0, adapt_when, reset_mean_at_adapt) for the three loops."""
    d = len(theta0)
    if variant == "fit":      # PyHillFit.py:748-751, 787
        return 0.05 * np.diag(np.abs(theta0)), 1000 * d, False
    if variant == "hier":     # PyHillFit.py:431, 440
        return np.diag(0.01 * np.abs(theta0)), 100 * d, False
    if variant == "temp":     # PyHillTemp.py:80, 83, 114-115
        return np.eye(d), 1000 * d, True
    raise ValueError(variant)


def adaptive_metropolis(target, theta0, iterations, thinning, variant, rng="numpy", seed=1, chain_id=0,
                        state_out=None):
    """One chain.  `target(theta) -> log target`.  Returns (chain[iterations//thinning + 1, d+1], acceptance).

    rng="numpy": theta* = npr.multivariate_normal(theta, exp(loga)*cov); u = npr.rand() drawn from the
    *global* numpy RandomState, exactly like the reference (caller seeds it: npr.seed(25) at
    PyHillFit.py:824-825, npr.seed(1) at PyHillTemp.py:16-17, nothing for the hierarchical loop).
    rng="philox": the CUDA kernels' stream (see above) with theta* = theta + exp(loga/2) chol(cov) z.
    """
    import numpy.random as npr
    theta_cur = np.array(theta0, dtype=float)
    d = len(theta_cur)
    cov, adapt_when, reset_mean = am_defaults(variant, theta_cur)
    if rng == "philox":
        dg = np.diag(cov).copy()
        cov[np.diag_indices(d)] = np.where(dg > 0, dg, COV0_DIAG_FLOOR)
    mean = np.copy(theta_cur)
    log_target_cur = target(theta_cur)
    num_saved = iterations // thinning + 1
    chain = np.zeros((num_saved, d + 1))
    chain[0, :] = np.concatenate((theta_cur, [log_target_cur]))
    loga = 0.
    acceptance = 0.
    t = 1
    while t <= iterations:
        if rng == "numpy":
            theta_star = npr.multivariate_normal(theta_cur, np.exp(loga) * cov)
            log_target_star = target(theta_star)
            u = npr.rand()
        else:
            u, z = philox_draw(seed, chain_id, t, d)
            theta_star = theta_cur + math.exp(0.5 * loga) * (guarded_cholesky(cov) @ z)
            log_target_star = target(theta_star)
        if np.log(u) < log_target_star - log_target_cur:
            theta_cur = theta_star
            log_target_cur = log_target_star
            accepted = 1
        else:
            accepted = 0
        acceptance = ((t - 1.) * acceptance + accepted) / t
        if t % thinning == 0:
            chain[t // thinning, :] = np.concatenate((theta_cur, [log_target_cur]))
        if reset_mean and t == adapt_when:
            mean = np.copy(theta_cur)
        if t > adapt_when:
            s = t - adapt_when
            gamma_s = 1. / (s + 1.) ** 0.6
            bit = np.array([theta_cur - mean])
            cov = (1 - gamma_s) * cov + gamma_s * np.dot(np.transpose(bit), bit)
            mean = (1 - gamma_s) * mean + gamma_s * theta_cur
            loga += gamma_s * (accepted - 0.25)
        t += 1
    if state_out is not None:
        state_out.update(theta=theta_cur, log_target=log_target_cur, mean=mean, cov=cov, loga=loga,
                         acceptance=acceptance)
    return chain, acceptance


# ----------------------------------------------------------------------------
# posterior-predictive CDFs -- python/construct_hierarchical_cdfs.py:32-58
# ----------------------------------------------------------------------------
def construct_posterior_predictive_cdfs(alphas, betas, mus, ss):
    num_x_pts = 501
    hill_x_range = np.linspace(0., 4., num_x_pts)
    pic50_x_range = np.linspace(-2., 12., num_x_pts)
    num_iterations = len(alphas)
    hill_pdf_sum = np.zeros(num_x_pts)
    hill_cdf_sum = np.zeros(num_x_pts)
    pic50_pdf_sum = np.zeros(num_x_pts)
    pic50_cdf_sum = np.zeros(num_x_pts)
    for i in range(num_iterations):
        hill_cdf_sum += st.fisk.cdf(hill_x_range, c=betas[i], scale=alphas[i], loc=0)
        hill_pdf_sum += st.fisk.pdf(hill_x_range, c=betas[i], scale=alphas[i], loc=0)
        pic50_cdf_sum += st.logistic.cdf(pic50_x_range, mus[i], ss[i])
        pic50_pdf_sum += st.logistic.pdf(pic50_x_range, mus[i], ss[i])
    return (hill_x_range, hill_cdf_sum / num_iterations, pic50_x_range, pic50_cdf_sum / num_iterations,
            hill_pdf_sum / num_iterations, pic50_pdf_sum / num_iterations)


# ----------------------------------------------------------------------------
# Speculative (prefetching) evaluation of the same loop -- an executable statement of the round protocol of
# pyhillfit_b200/csrc/phf_single_spec.cu, used by tests/test_oracle_golden.py to show that it commits exactly the
# sequential chain: S proposals per round, proposal g made from the state "the g iterations before it were rejected"
# (theta and the log-target unchanged, mean / covariance / loga adapted with accepted = 0), the first accepted one ends
# the round, the state after k-1 rejections is hypothesis k-1 itself.
# ----------------------------------------------------------------------------
def adaptive_metropolis_speculative(target, theta0, iterations, thinning, variant, depth, seed=1, chain_id=0):
    """Philox stream only.  Returns (chain, acceptance, rounds) -- chain and acceptance equal to
    adaptive_metropolis(..., rng="philox") bit for bit; `rounds` is how many rounds the iterations took."""
    theta = np.array(theta0, dtype=float)
    d = len(theta)
    cov, adapt_when, reset_mean = am_defaults(variant, theta)
    dg = np.diag(cov).copy()
    cov[np.diag_indices(d)] = np.where(dg > 0, dg, COV0_DIAG_FLOOR)
    mean = np.copy(theta)
    lt = target(theta)
    chain = np.zeros((iterations // thinning + 1, d + 1))
    chain[0, :] = np.concatenate((theta, [lt]))
    loga = 0.
    n_acc = 0
    t = 0          # iterations completed
    rounds = 0

    def adapt(mean, cov, loga, th, ti, accepted):
        # the adaptation after iteration ti (PyHillFit.py:840-846, PyHillTemp.py:114-122), th already post-accept
        if reset_mean and ti == adapt_when:
            mean = np.copy(th)
        if ti > adapt_when:
            gamma_s = 1. / ((ti - adapt_when) + 1.) ** 0.6
            bit = np.array([th - mean])
            cov = (1 - gamma_s) * cov + gamma_s * np.dot(np.transpose(bit), bit)
            mean = (1 - gamma_s) * mean + gamma_s * th
            loga = loga + gamma_s * (accepted - 0.25)
        return mean, cov, loga

    while t < iterations:
        rounds += 1
        n_valid = min(depth, iterations - t)
        hyp, props = [], []
        h_mean, h_cov, h_loga = mean, cov, loga
        for g in range(n_valid):                      # hypothesis g: iterations t+1 .. t+g rejected
            if g > 0:
                h_mean, h_cov, h_loga = adapt(h_mean, h_cov, h_loga, theta, t + g, 0)
            u, z = philox_draw(seed, chain_id, t + 1 + g, d)
            star = theta + math.exp(0.5 * h_loga) * (guarded_cholesky(h_cov) @ z)
            hyp.append((h_mean, h_cov, h_loga))
            props.append((u, star, target(star)))     # (every group evaluates its target in parallel on the device)
        first = next((g for g in range(n_valid) if np.log(props[g][0]) < props[g][2] - lt), None)
        k = n_valid if first is None else first + 1
        for j in range(1, k):                         # rows that fall on the rejected iterations: state unchanged
            if (t + j) % thinning == 0:
                chain[(t + j) // thinning, :] = np.concatenate((theta, [lt]))
        mean, cov, loga = hyp[k - 1]                  # the state after k-1 rejections IS hypothesis k-1
        accepted = 0
        if first is not None:
            theta, lt = props[first][1], props[first][2]
            accepted = 1
            n_acc += 1
        mean, cov, loga = adapt(mean, cov, loga, theta, t + k, accepted)
        if (t + k) % thinning == 0:
            chain[(t + k) // thinning, :] = np.concatenate((theta, [lt]))
        t += k
    return chain, n_acc / float(iterations), rounds
