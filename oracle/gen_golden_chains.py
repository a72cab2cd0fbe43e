"""Reference MCMC chains -> tests/golden/ref_chains.npz (summaries only; run via gen_golden.py chains).

* tempered single-level chains: the reference's own `do_mcmc` (python/PyHillTemp.py:57-125) extracted by
  oracle/ref_shim.py, numpy MT19937 + SVD proposals, theta0 = ones, run for Amiodarone/hERG on the
  reference's 41-point ladder for models 1 and 2 (-> ln B12 exactly as compute_bayes_factors.py does),
  plus three other pairs at T=1.
* PyHillFit-variant single-level chains (python/PyHillFit.py:748-751, 787-856: Sigma0 = 0.05 diag|theta0|, no mean
  reset, npr.seed(25)): that loop is inline script code behind `import cma`, so -- exactly like the hierarchical
  loop below -- the oracle's restatement of the loop (variant "fit", numpy RNG) drives the reference's OWN
  `dr.log_target`; theta0 = the repository's least-squares start (stored in the fixture; the reference's CMA-ES is
  not installed), three pairs (plain / two responses at 100 / the dropped -2.6 response) x both models.
* hierarchical: the reference's `log_target_distribution` (python/PyHillFit.py:173-193) driven by the
  oracle's restatement of the loop at PyHillFit.py:481-511 (the reference loop is inline script code
  that needs `cma`, so it cannot be executed as is).
"""
import multiprocessing as mp
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402
from pyhillfit_b200.ess import ess_geyer, ess_quantile_indicator  # noqa: E402

QS = [5, 25, 50, 75, 95]
ITERS = int(os.environ.get("PHF_GOLD_ITERS", 100000))
THIN = 5
BURN_FRAC = 4


def _pair(dr, drug, channel):
    num_expts, _, experiments = dr.load_crumb_data(drug, channel)
    concs = np.concatenate([experiments[i][:, 0] for i in range(num_expts)])
    responses = np.concatenate([experiments[i][:, 1] for i in range(num_expts)])
    return experiments, concs, responses


def _summ(chain, d):
    q = np.percentile(chain[:, :d], QS, axis=0)          # [5, d]
    mean = chain[:, :d].mean(axis=0)
    sd = chain[:, :d].std(axis=0, ddof=1)
    ess = np.array([ess_geyer(chain[:, j]) for j in range(d)])
    # ESS of the indicator 1[x <= q_p] at each recorded quantile: the Monte-Carlo error of the quantile ESTIMATE is
    # sqrt(p(1-p)/ess_q) / density (pyhillfit_b200/ess.py); the tests take the density from the pooled GPU sample
    ess_q = np.array([[ess_quantile_indicator(chain[:, j], q[i, j]) for j in range(d)] for i in range(len(QS))])
    return q, mean, sd, ess, ess_q


def _run_temp(job):
    drug, channel, model, temperature = job
    import io, contextlib
    dr = ref_shim.load_doseresponse()
    dr.setup(os.path.join(ref_shim.REF_ROOT, "data", "crumb_data.csv"))
    dr.define_model(model)
    _, concs, responses = _pair(dr, drug, channel)
    ns = dict(args=types.SimpleNamespace(iterations=ITERS, thinning=THIN, burn_in_fraction=BURN_FRAC),
              responses=responses, where_r_0=responses == 0, where_r_100=responses == 100,
              where_r_other=(0 < responses) & (responses < 100), concs=concs, num_params=dr.num_params)
    ns["pi_bit"] = dr.compute_pi_bit_of_log_likelihood(ns["where_r_other"])
    do_mcmc, _ = ref_shim.load_do_mcmc(dr, ns)
    import numpy.random as npr
    npr.seed(1)  # PyHillTemp.py:16-17 (forked workers inherit this state)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        chain = do_mcmc(temperature)
    d = dr.num_params
    with np.errstate(all="ignore"):
        ll1 = np.array([dr.log_data_likelihood(responses, ns["where_r_0"], ns["where_r_100"], ns["where_r_other"],
                                               concs, chain[i, :d], 1, ns["pi_bit"]) for i in range(len(chain))])
    q, mean, sd, ess, ess_q = _summ(chain, d)
    return dict(q=q, mean=mean, sd=sd, ess=ess, ess_q=ess_q, ll1_mean=ll1.mean(), ll1_sd=ll1.std(ddof=1),
                ll1_ess=ess_geyer(ll1), rows=len(chain))


def _run_fit(job):
    """PyHillFit-variant chain: oracle loop (variant "fit", numpy RNG, npr.seed(25) as at PyHillFit.py:824-825) around
    the reference's own dr.log_target (doseresponse.py:187-189)."""
    drug, channel, model, iters = job
    import hill_oracle as ho
    import numpy.random as npr
    from pyhillfit_b200.initial_fit import best_fit
    dr = ref_shim.load_doseresponse()
    dr.setup(os.path.join(ref_shim.REF_ROOT, "data", "crumb_data.csv"))
    dr.define_model(model)
    _, concs, responses = _pair(dr, drug, channel)
    w0, w100, wo = responses == 0, responses == 100, (0 < responses) & (responses < 100)
    pi_bit = dr.compute_pi_bit_of_log_likelihood(wo)
    theta0, _ = best_fit(model, concs, responses)

    def target(th):
        with np.errstate(all="ignore"):
            return dr.log_target(responses, w0, w100, wo, concs, th, 1, pi_bit)

    npr.seed(25)
    chain, acc = ho.adaptive_metropolis(target, theta0, iters, THIN, "fit", rng="numpy")
    chain = chain[len(chain) // BURN_FRAC:]    # PyHillFit.py:861-864
    d = dr.num_params
    q, mean, sd, ess, ess_q = _summ(chain, d)
    return dict(q=q, mean=mean, sd=sd, ess=ess, ess_q=ess_q, theta0=np.asarray(theta0, dtype=float), acc=acc,
                rows=len(chain), lt_mean=chain[:, d].mean())


def _run_hier(job):
    drug, channel, iters = job
    import hill_oracle as ho
    import numpy.random as npr
    dr = ref_shim.load_doseresponse()
    dr.setup(os.path.join(ref_shim.REF_ROOT, "data", "crumb_data.csv"))
    experiments, _, _ = _pair(dr, drug, channel)
    h = ref_shim.load_hierarchical_functions(dr)
    shapes, scales, locs = ho.hier_prior_constants()
    ne = len(experiments)
    theta0 = np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([6.0, 1.0], ne), [6.0]))

    def target(th):
        with np.errstate(all="ignore"):
            return h["log_target_distribution"](experiments, th, shapes, scales, locs)

    npr.seed(7)
    chain, acc = ho.adaptive_metropolis(target, theta0, iters, THIN, "hier", rng="numpy")
    chain = chain[len(chain) // BURN_FRAC:]
    d = len(theta0)
    q, mean, sd, ess, ess_q = _summ(chain, d)
    return dict(q=q, mean=mean, sd=sd, ess=ess, ess_q=ess_q, theta0=theta0, acc=acc, rows=len(chain))


LADDER_PAIRS = [("Amiodarone", "hERG", "ladder"), ("Bepridil", "hERG", "ladder2")]   # ladder2: two responses at 100
FIT_PAIRS = [("Amiodarone", "hERG"), ("Bepridil", "hERG"), ("Amitriptyline", "Kv4.3")]


def _key(drug, channel):
    return "%s_%s" % (drug, channel.replace(".", "_"))


def main(dr):
    gold = os.path.join(ROOT, "tests", "golden")
    temps = (np.arange(dr.n + 1.) / dr.n) ** dr.c
    jobs = [(dg, ch, m, float(t)) for dg, ch, _ in LADDER_PAIRS for m in (1, 2) for t in temps]
    extra = [("Bepridil", "hERG", 2, 1.0), ("Amitriptyline", "Kv4.3", 2, 1.0), ("Bepridil", "hERG", 1, 1.0)]
    hier_jobs = [("Amiodarone", "hERG", 2 * ITERS), ("Dofetilide", "hERG", 2 * ITERS)]
    fit_jobs = [(dg, ch, m, 2 * ITERS) for dg, ch in FIT_PAIRS for m in (1, 2)]
    with mp.Pool(min(8, mp.cpu_count())) as pool:
        r_h = pool.map_async(_run_hier, hier_jobs)
        r_f = pool.map_async(_run_fit, fit_jobs, chunksize=1)
        r_t = pool.map_async(_run_temp, jobs + extra, chunksize=1)
        res_t = r_t.get()
        res_f = r_f.get()
        res_h = r_h.get()
    out = {"temps": temps, "iters": ITERS, "thin": THIN, "burn_frac": BURN_FRAC, "quantiles": np.array(QS)}
    for dg, ch, tag in LADDER_PAIRS:
        for m in (1, 2):
            rs = [r for j, r in zip(jobs, res_t[:len(jobs)]) if j[0] == dg and j[1] == ch and j[2] == m]
            for f in ("q", "mean", "sd", "ess", "ess_q"):
                out["%s_m%d_%s" % (tag, m, f)] = np.stack([r[f] for r in rs])
            for f in ("ll1_mean", "ll1_sd", "ll1_ess"):
                out["%s_m%d_%s" % (tag, m, f)] = np.array([r[f] for r in rs])
        lp = [0.5 * np.sum((temps[1:] - temps[:-1]) * (out["%s_m%d_ll1_mean" % (tag, m)][1:] +
                                                         out["%s_m%d_ll1_mean" % (tag, m)][:-1])) for m in (1, 2)]
        sfx = "" if tag == "ladder" else "_" + tag
        out["log_py_m1" + sfx], out["log_py_m2" + sfx], out["B12" + sfx] = lp[0], lp[1], np.exp(lp[0] - lp[1])
        print("%s/%s: log p(y|M1)=%.4f log p(y|M2)=%.4f B12=%.4g" % (dg, ch, lp[0], lp[1], np.exp(lp[0] - lp[1])))
    out["ladder2_pair"] = np.array(LADDER_PAIRS[1][:2])
    for (drug, channel, m, t), r in zip(extra, res_t[len(jobs):]):
        k = "extra_%s_m%d" % (_key(drug, channel), m)
        for f in ("q", "mean", "sd", "ess", "ess_q", "ll1_mean", "ll1_sd", "ll1_ess"):
            out[k + "_" + f] = r[f]
    for (drug, channel, m, iters), r in zip(fit_jobs, res_f):
        k = "fit_%s_m%d" % (_key(drug, channel), m)
        for f in ("q", "mean", "sd", "ess", "ess_q", "theta0", "acc", "lt_mean"):
            out[k + "_" + f] = r[f]
        out[k + "_iters"] = iters
    for (drug, channel, iters), r in zip(hier_jobs, res_h):
        k = "hier_%s_%s" % (drug, channel)
        for f in ("q", "mean", "sd", "ess", "ess_q", "theta0"):
            out[k + "_" + f] = r[f]
        out[k + "_iters"] = iters
        out[k + "_acc"] = r["acc"]
    np.savez_compressed(os.path.join(gold, "ref_chains.npz"), **out)


if __name__ == "__main__":
    d = ref_shim.load_doseresponse()
    main(d)
