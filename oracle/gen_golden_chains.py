"""Reference MCMC chains -> tests/golden/ref_chains.npz (summaries only; run via gen_golden.py chains).

* tempered single-level chains: the reference's own `do_mcmc` (python/PyHillTemp.py:57-125) extracted by
  oracle/ref_shim.py, numpy MT19937 + SVD proposals, theta0 = ones, run for Amiodarone/hERG on the
  reference's 41-point ladder for models 1 and 2 (-> ln B12 exactly as compute_bayes_factors.py does),
  plus three other pairs at T=1.
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
from pyhillfit_b200.ess import ess_geyer  # noqa: E402

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
    return q, mean, sd, ess


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
    q, mean, sd, ess = _summ(chain, d)
    return dict(q=q, mean=mean, sd=sd, ess=ess, ll1_mean=ll1.mean(), ll1_sd=ll1.std(ddof=1), ll1_ess=ess_geyer(ll1),
                rows=len(chain))


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
    q, mean, sd, ess = _summ(chain, d)
    return dict(q=q, mean=mean, sd=sd, ess=ess, theta0=theta0, acc=acc, rows=len(chain))


def main(dr):
    gold = os.path.join(ROOT, "tests", "golden")
    temps = (np.arange(dr.n + 1.) / dr.n) ** dr.c
    jobs = [("Amiodarone", "hERG", m, float(t)) for m in (1, 2) for t in temps]
    extra = [("Bepridil", "hERG", 2, 1.0), ("Amitriptyline", "Kv4.3", 2, 1.0), ("Bepridil", "hERG", 1, 1.0)]
    hier_jobs = [("Amiodarone", "hERG", 2 * ITERS), ("Dofetilide", "hERG", 2 * ITERS)]
    with mp.Pool(min(8, mp.cpu_count())) as pool:
        r_h = pool.map_async(_run_hier, hier_jobs)
        r_t = pool.map_async(_run_temp, jobs + extra, chunksize=1)
        res_t = r_t.get()
        res_h = r_h.get()
    out = {"temps": temps, "iters": ITERS, "thin": THIN, "burn_frac": BURN_FRAC, "quantiles": np.array(QS)}
    for m in (1, 2):
        rs = [r for j, r in zip(jobs, res_t[:len(jobs)]) if j[2] == m]
        d = 2 if m == 1 else 3
        out["ladder_m%d_q" % m] = np.stack([r["q"] for r in rs])
        out["ladder_m%d_mean" % m] = np.stack([r["mean"] for r in rs])
        out["ladder_m%d_sd" % m] = np.stack([r["sd"] for r in rs])
        out["ladder_m%d_ess" % m] = np.stack([r["ess"] for r in rs])
        out["ladder_m%d_ll1_mean" % m] = np.array([r["ll1_mean"] for r in rs])
        out["ladder_m%d_ll1_sd" % m] = np.array([r["ll1_sd"] for r in rs])
        out["ladder_m%d_ll1_ess" % m] = np.array([r["ll1_ess"] for r in rs])
    for (drug, channel, m, t), r in zip(extra, res_t[len(jobs):]):
        k = "extra_%s_%s_m%d" % (drug, channel.replace(".", "_"), m)
        for f in ("q", "mean", "sd", "ess"):
            out[k + "_" + f] = r[f]
        out[k + "_ll1_mean"] = r["ll1_mean"]
    for (drug, channel, iters), r in zip(hier_jobs, res_h):
        k = "hier_%s_%s" % (drug, channel)
        for f in ("q", "mean", "sd", "ess", "theta0"):
            out[k + "_" + f] = r[f]
        out[k + "_iters"] = iters
        out[k + "_acc"] = r["acc"]
    lp1 = 0.5 * np.sum((temps[1:] - temps[:-1]) * (out["ladder_m1_ll1_mean"][1:] + out["ladder_m1_ll1_mean"][:-1]))
    lp2 = 0.5 * np.sum((temps[1:] - temps[:-1]) * (out["ladder_m2_ll1_mean"][1:] + out["ladder_m2_ll1_mean"][:-1]))
    out["log_py_m1"], out["log_py_m2"], out["B12"] = lp1, lp2, np.exp(lp1 - lp2)
    np.savez_compressed(os.path.join(gold, "ref_chains.npz"), **out)
    print("ref_chains.npz: log p(y|M1)=%.4f log p(y|M2)=%.4f B12=%.4g" % (lp1, lp2, out["B12"]))


if __name__ == "__main__":
    d = ref_shim.load_doseresponse()
    main(d)
