"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (via oracle/ref_shim.py).

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py data targets        # seconds
    python oracle/gen_golden.py chains              # minutes, uses all cores
Fixtures are committed; the GPU box never runs this.
"""
import itertools as it
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

DATA = os.path.join(ref_shim.REF_ROOT, "data")


def _load(dr, name):
    dr.setup(os.path.join(DATA, name + ".csv"))
    return dr.df


def gen_data(dr):
    """The three input CSVs as column arrays (inputs, not code)."""
    out = {}
    for name in ["crumb_data", "modified_crumb_data", "synthetic_data"]:
        df = _load(dr, name)
        out[name + "__drug"] = np.array(df.Drug.astype(str).to_numpy(), dtype="U")
        out[name + "__channel"] = np.array(df.Channel.astype(str).to_numpy(), dtype="U")
        out[name + "__experiment"] = df.Experiment.to_numpy().astype(np.int32)
        out[name + "__dose"] = df.Concentration.to_numpy().astype(np.float64)
        out[name + "__response"] = df.Inhibition.to_numpy().astype(np.float64)
    np.savez_compressed(os.path.join(GOLD, "datasets.npz"), **out)
    print("datasets.npz", {k: v.shape for k, v in out.items() if k.endswith("dose")})


def _pair_arrays(dr, drug, channel):
    num_expts, experiment_numbers, experiments = dr.load_crumb_data(drug, channel)
    concs = np.array([])
    responses = np.array([])
    for i in range(num_expts):  # PyHillFit.py:661-665
        concs = np.concatenate((concs, experiments[i][:, 0]))
        responses = np.concatenate((responses, experiments[i][:, 1]))
    w0 = responses == 0
    w100 = responses == 100
    wo = (0 < responses) & (responses < 100)
    return experiments, concs, responses, w0, w100, wo


TEMPS_CYCLE = np.array([1.0, 0.125, 0.0, (1 / 40.) ** 3, (27 / 40.) ** 3, 1.0, 0.5, (39 / 40.) ** 3])


def _thetas(rng, model, nt):
    pic50 = rng.uniform(-4, 12, nt)
    hill = rng.uniform(-0.4, 10.4, nt)
    sigma = np.exp(rng.uniform(np.log(6e-4), np.log(60.), nt))
    # a band of plausible posteriors-region values so that not everything is deep in the tails
    k = nt // 2
    pic50[:k] = rng.uniform(3, 9, k)
    hill[:k] = rng.uniform(0.2, 3, k)
    sigma[:k] = rng.uniform(1, 15, k)
    th = np.stack([pic50, hill, sigma], axis=1)
    # hand-picked edge cases
    edge = np.array([[6, 1, 5.], [5.5, 0.8, 8.], [1, 1, 1.], [-3.5, 11, 5e-4], [6, 1, 1e-3], [400, 1, 5.],
                     [-3, 0, 1.0010001e-3], [6, 10, 50.], [6, 0, 5.], [-3.0000001, 1, 5], [12, 9.99, 0.0011],
                     [330, 2, 3.]])
    th[-len(edge):] = edge
    return th if model == 2 else th[:, [0, 2]]


def gen_targets(dr):
    df = _load(dr, "crumb_data")
    pairs = list(it.product(dr.drugs, dr.channels))
    rng = np.random.default_rng(20161018)
    NT = 48
    out = {"pairs_drug": np.array([p[0] for p in pairs], dtype="U"),
           "pairs_channel": np.array([p[1] for p in pairs], dtype="U")}
    pi_bits = np.zeros(len(pairs))
    for model in (1, 2):
        dr.define_model(model)
        d = dr.num_params
        TH = np.zeros((len(pairs), NT, d))
        TT = np.zeros((len(pairs), NT))
        LT = np.zeros((len(pairs), NT))
        LL = np.zeros((len(pairs), NT))
        LL1 = np.zeros((len(pairs), NT))
        for ip, (drug, channel) in enumerate(pairs):
            _, concs, y, w0, w100, wo = _pair_arrays(dr, drug, channel)
            pb = dr.compute_pi_bit_of_log_likelihood(wo)
            pi_bits[ip] = pb
            th = _thetas(rng, model, NT)
            tt = TEMPS_CYCLE[(np.arange(NT) + ip) % len(TEMPS_CYCLE)]
            for k in range(NT):
                LT[ip, k] = dr.log_target(y, w0, w100, wo, concs, th[k], tt[k], pb)
                LL[ip, k] = dr.log_data_likelihood(y, w0, w100, wo, concs, th[k], tt[k], pb)
                LL1[ip, k] = dr.log_data_likelihood(y, w0, w100, wo, concs, th[k], 1, pb)
            TH[ip], TT[ip] = th, tt
        out["theta_m%d" % model] = TH
        out["t_m%d" % model] = TT
        out["log_target_m%d" % model] = LT
        out["log_lik_m%d" % model] = LL
        out["log_lik_t1_m%d" % model] = LL1
    out["pi_bit"] = pi_bits
    np.savez_compressed(os.path.join(GOLD, "log_target_golden.npz"), **out)
    print("log_target_golden.npz", out["log_target_m2"].shape,
          "finite frac", np.isfinite(out["log_target_m2"]).mean())

    # hierarchical target (PyHillFit.py:173-193) on every pair
    h = ref_shim.load_hierarchical_functions(dr)
    import hill_oracle as ho
    shapes, scales, locs = ho.hier_prior_constants()
    NTH = 24
    DMAX = 17
    TH = np.full((len(pairs), NTH, DMAX), np.nan)
    LT = np.zeros((len(pairs), NTH))
    NE = np.zeros(len(pairs), dtype=np.int32)
    for ip, (drug, channel) in enumerate(pairs):
        experiments, *_ = _pair_arrays(dr, drug, channel)
        ne = len(experiments)
        NE[ip] = ne
        dim = 5 + 2 * ne
        for k in range(NTH):
            th = np.zeros(dim)
            th[0] = rng.uniform(0.05, 3)
            th[1] = rng.uniform(2.01, 12)
            th[2] = rng.uniform(-3.5, 10)
            th[3] = rng.uniform(0.011, 2)
            th[4:-1:2] = rng.uniform(-1.9, 10, ne)
            th[5:-1:2] = rng.uniform(0.0, 5, ne)
            th[-1] = np.exp(rng.uniform(np.log(0.05), np.log(40.)))
            if k % 8 == 7:  # out-of-support or boundary cases, one condition at a time
                which = (k // 8 + ip) % 7
                if which == 0: th[1] = 2.0
                elif which == 1: th[3] = 0.01
                elif which == 2: th[5] = -1e-9
                elif which == 3: th[4] = -2.0000001
                elif which == 4: th[-1] = 1e-3
                elif which == 5: th[5] = 0.0
                elif which == 6: th[2] = 900.; th[3] = 0.011   # exp(-z) overflow artefact -> -inf (PyHillFit.py:145-146)
            with np.errstate(all="ignore"):
                LT[ip, k] = h["log_target_distribution"](experiments, th, shapes, scales, locs)
            TH[ip, k, :dim] = th
    np.savez_compressed(os.path.join(GOLD, "hier_target_golden.npz"), theta=TH, log_target=LT, ne=NE,
                        shapes=shapes, scales=scales, locs=locs,
                        pairs_drug=out["pairs_drug"], pairs_channel=out["pairs_channel"])
    print("hier_target_golden.npz", LT.shape, "finite frac", np.isfinite(LT).mean(), "nan", np.isnan(LT).sum())

    # chaste/samples: shipped (alpha, mu) posterior draws -> per-pair summary (soft fixture)
    summ = np.full((len(pairs), 4), np.nan)
    for ip, (drug, channel) in enumerate(pairs):
        f = os.path.join(ref_shim.REF_ROOT, "chaste", "samples",
                         "%s_%s_hill_pic50_samples.txt" % (drug.replace("/", "_"), channel.replace("/", "_")))
        if os.path.exists(f):
            a = np.loadtxt(f)
            summ[ip] = [a[:, 0].mean(), a[:, 0].std(), a[:, 1].mean(), a[:, 1].std()]
    np.savez_compressed(os.path.join(GOLD, "chaste_alpha_mu_summary.npz"), summary=summ,
                        pairs_drug=out["pairs_drug"], pairs_channel=out["pairs_channel"])
    print("chaste summary: pairs with samples", np.isfinite(summ[:, 0]).sum())


if __name__ == "__main__":
    what = sys.argv[1:] or ["data", "targets"]
    os.makedirs(GOLD, exist_ok=True)
    dr = ref_shim.load_doseresponse()
    if "data" in what:
        gen_data(dr)
    if "targets" in what:
        gen_targets(dr)
    if "chains" in what:
        import gen_golden_chains
        gen_golden_chains.main(dr)
