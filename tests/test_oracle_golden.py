"""Pin both oracles (numpy/scipy restatement and C restatement) to outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import c_oracle
import hill_oracle as ho
from _data import GOLD, Table

TOL = 1e-13  # |a-b| <= TOL * max(1, |b|); +-inf must match exactly


def assert_close(a, b, tol=TOL):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    inf = ~np.isfinite(b)
    assert np.array_equal(a[inf], b[inf]), "non-finite values differ"
    err = np.abs(a[~inf] - b[~inf]) / np.maximum(1.0, np.abs(b[~inf]))
    assert err.size == 0 or err.max() <= tol, "max scaled error %.3e" % err.max()


@pytest.fixture(scope="module")
def table():
    return Table("crumb_data")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "log_target_golden.npz"))


def test_pairs_order_matches_reference(table, gold):
    pairs = table.pairs()
    assert len(pairs) == 210
    assert [p[0] for p in pairs] == list(gold["pairs_drug"])
    assert [p[1] for p in pairs] == list(gold["pairs_channel"])


@pytest.mark.parametrize("model", [1, 2])
def test_c_oracle_log_target_all_pairs(table, gold, model):
    for ip, (drug, channel) in enumerate(table.pairs()):
        concs, y = table.concat(drug, channel)
        pb = ho.compute_pi_bit_of_log_likelihood(y)
        assert pb == gold["pi_bit"][ip]
        th = gold["theta_m%d" % model][ip]
        tt = gold["t_m%d" % model][ip]
        lt, ll1 = c_oracle.log_target_batch(model, concs, y, th, tt, pb)
        assert_close(lt, gold["log_target_m%d" % model][ip])
        assert_close(ll1, gold["log_lik_t1_m%d" % model][ip])


@pytest.mark.parametrize("model", [1, 2])
def test_py_oracle_log_target_subset(table, gold, model):
    pairs = table.pairs()
    for ip in list(range(0, 210, 9)) + [12, 40]:  # 12: Amitriptyline/Kv4.3 (negative response)
        drug, channel = pairs[ip]
        concs, y = table.concat(drug, channel)
        w0, w100, wo = ho.masks(y)
        pb = ho.compute_pi_bit_of_log_likelihood(wo)
        th = gold["theta_m%d" % model][ip]
        tt = gold["t_m%d" % model][ip]
        with np.errstate(all="ignore"):
            lt = [ho.log_target(model, y, w0, w100, wo, concs, th[k], tt[k], pb) for k in range(len(th))]
            ll = [ho.log_data_likelihood(model, y, w0, w100, wo, concs, th[k], tt[k], pb) for k in range(len(th))]
        # same numpy/scipy calls as the reference -> bit-identical
        assert np.array_equal(np.array(lt, dtype=float), gold["log_target_m%d" % model][ip])
        assert np.array_equal(np.array(ll, dtype=float), gold["log_lik_m%d" % model][ip])


def test_survey_known_answers(table):
    """SURVEY.md section 8c values (Amiodarone/hERG etc.), both oracles."""
    concs, y = table.concat("Amiodarone", "hERG")
    w0, w100, wo = ho.masks(y)
    pb = ho.compute_pi_bit_of_log_likelihood(wo)
    assert pb == 11.027262398456072
    cases = [((6, 1, 5.), 1, -58.39140921642633), ((6, 1, 5.), 0.125, -5.633162956781316),
             ((6, 1, 5.), 0, 1.9037293660251158), ((5.5, 0.8, 8.), 1, -55.87532519931496),
             ((1, 1, 1.), 1, -12946.405169647265), ((400, 1, 5.), 1, -1331.5327403576296)]
    for th, t, want in cases:
        with np.errstate(all="ignore"):
            got = ho.log_target(2, y, w0, w100, wo, concs, np.array(th, dtype=float), t, pb)
        assert got == pytest.approx(want, rel=1e-14)
        got_c, _ = c_oracle.log_target_batch(2, concs, y, np.array([th], dtype=float), t, pb)
        assert got_c[0] == pytest.approx(want, rel=1e-13)
    got_c, _ = c_oracle.log_target_batch(1, concs, y, np.array([[5.5, 8.]]), 1, pb)
    assert got_c[0] == pytest.approx(-63.44284177305924, rel=1e-13)
    concs, y = table.concat("Amitriptyline", "Kv4.3")
    assert len(y) == 19 and (y < 0).sum() == 1
    pb = ho.compute_pi_bit_of_log_likelihood(y)
    assert pb == pytest.approx(17.45983213088878, rel=1e-15)
    got_c, _ = c_oracle.log_target_batch(2, concs, y, np.array([[5.5, 0.8, 8.]]), 1, pb)
    assert got_c[0] == pytest.approx(-133.0008537658694, rel=1e-13)


def test_hier_oracles(table):
    g = np.load(os.path.join(GOLD, "hier_target_golden.npz"))
    shapes, scales, locs = ho.hier_prior_constants()
    assert np.array_equal(shapes, g["shapes"]) and np.array_equal(scales, g["scales"]) and np.array_equal(locs, g["locs"])
    for ip, (drug, channel) in enumerate(table.pairs()):
        ex = table.experiments(drug, channel)
        ne = len(ex)
        assert ne == g["ne"][ip]
        th = np.ascontiguousarray(g["theta"][ip][:, :5 + 2 * ne])
        got = c_oracle.hier_log_target_batch(ex, th, shapes, scales, locs)
        assert_close(got, g["log_target"][ip], 2e-13)
        if ip % 15 == 0:
            with np.errstate(all="ignore"):
                got_py = np.array([ho.hier_log_target(ex, t, shapes, scales, locs) for t in th], dtype=float)
            assert np.array_equal(got_py, g["log_target"][ip])
    assert ho.hier_log_target(table.experiments("Amiodarone", "hERG"),
                              np.array([1.2, 6, 5.9, .15, 6, 1, 6.1, .9, 5.9, 1.1, 5]), shapes, scales, locs) \
        == pytest.approx(-45.839513564246495, rel=1e-14)


def test_log_ndtr_against_scipy():
    from scipy.special import log_ndtr, ndtr
    x = np.concatenate([np.linspace(-60, 8, 4001), -np.logspace(-8, 5, 500)])
    L = c_oracle.lib()
    got = np.array([L.phf_oracle_log_ndtr(v) for v in x])
    assert_close(got, log_ndtr(x), 5e-15)
    got = np.array([L.phf_oracle_ndtr(v) for v in np.linspace(-12, 12, 2001)])
    assert np.max(np.abs(got - ndtr(np.linspace(-12, 12, 2001)))) < 3e-16


def test_philox_known_answer_and_py_c_agree():
    # Random123 known-answer vectors for philox4x32-10
    assert ho.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert ho.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert ho.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    L = c_oracle.lib()
    w = np.zeros(4, dtype=np.uint32)
    for seed, chain, t, j in [(1, 0, 1, 0), (0xdeadbeefcafe, 2 ** 40 + 7, 499999, 3)]:
        L.phf_oracle_philox(seed, chain, t, j, w)
        assert tuple(int(v) for v in w) == ho.philox_call(seed, chain, t, j)
    for d in (2, 3, 11, 17):
        u = np.zeros(1)
        z = np.zeros(d + 1)
        L.phf_oracle_draw(12345, 77, 42, d, u, z)
        pu, pz = ho.philox_draw(12345, 77, 42, d)
        assert u[0] == pu
        assert np.allclose(z[:d], pz, rtol=0, atol=1e-15)


def test_am_c_oracle_follows_py_oracle(table):
    """Same Philox stream, same algorithm -> trajectories agree (C vs numpy restatement)."""
    concs, y = table.concat("Amiodarone", "hERG")
    w0, w100, wo = ho.masks(y)
    pb = ho.compute_pi_bit_of_log_likelihood(wo)
    for variant, model, temp, theta0 in [("temp", 2, 0.421875, np.ones(3)), ("fit", 2, 1, np.array([6.0, 0.9, 7.0])),
                                         ("fit", 1, 1, np.array([6.0, 7.0]))]:
        cov0, adapt_when, reset = ho.am_defaults(variant, theta0)
        adapt_when = 60  # exercise adaptation quickly
        iters, thin = 400, 5

        def target(th):
            with np.errstate(all="ignore"):
                return ho.log_target(model, y, w0, w100, wo, concs, th, temp, pb)

        # python loop with patched adapt_when
        orig = ho.am_defaults
        ho.am_defaults = lambda v, t0: (cov0, adapt_when, reset)
        try:
            chain_py, acc = ho.adaptive_metropolis(target, theta0, iters, thin, variant, rng="philox", seed=9,
                                                   chain_id=3)
        finally:
            ho.am_defaults = orig
        lt0, ll10 = c_oracle.log_target_batch(model, concs, y, theta0[None, :], temp, pb)
        st = c_oracle.make_state(theta0, lt0[0], ll10[0], cov0)
        chain_c = c_oracle.am_single(model, concs, y, temp, pb, st, 0, iters, thin, adapt_when, reset, 9, 3)
        assert np.allclose(chain_c, chain_py[1:], rtol=1e-9, atol=1e-9)
        assert st[-1] / iters == pytest.approx(acc, abs=1e-12)


def test_cdf_oracle_matches_reference_golden():
    """oracle restatement of construct_posterior_predictive_cdfs vs the unmodified reference (cdf_golden.npz)."""
    g = np.load(os.path.join(GOLD, "cdf_golden.npz"))
    r = g["rows"][:500]
    with np.errstate(all="ignore"):
        hx, hc, px, pc, hp, pp = ho.construct_posterior_predictive_cdfs(r[:, 0], r[:, 1], r[:, 2], r[:, 3])
    assert np.array_equal(hx, g["hill_x"]) and np.array_equal(px, g["pic50_x"])
    full = g["rows"]
    with np.errstate(all="ignore"):
        out = ho.construct_posterior_predictive_cdfs(full[:, 0], full[:, 1], full[:, 2], full[:, 3])
    for got, key in zip(out, ("hill_x", "hill_cdf", "pic50_x", "pic50_cdf", "hill_pdf", "pic50_pdf")):
        assert np.array_equal(got, g[key]), key


def test_am_hier_c_oracle_follows_py_oracle(table):
    """The C hierarchical AM loop (oracle/hill_oracle.c: what every GPU hierarchical trajectory test is compared with)
    is pinned to the numpy restatement of python/PyHillFit.py:431-511 driven by the restated target of
    python/PyHillFit.py:173-193 (itself pinned to the unmodified reference by hier_target_golden.npz), on the same
    Philox stream: Ne = 3 (dim 11) and Ne = 5 (dim 15), across the start of adaptation."""
    shapes, scales, locs = ho.hier_prior_constants()
    for drug, channel in (("Amiodarone", "hERG"), ("Dofetilide", "hERG")):
        ex = table.experiments(drug, channel)
        ne = len(ex)
        theta0 = np.concatenate(([1.1, 4.2, 6.0, 0.3], np.tile([6.0, 1.0], ne) + 0.05 * np.arange(2 * ne), [7.0]))
        cov0, _, reset = ho.am_defaults("hier", theta0)
        adapt_when, iters, thin, seed, cid = 40, 300, 5, 11, 5

        def target(th):
            with np.errstate(all="ignore"):
                return ho.hier_log_target(ex, th, shapes, scales, locs)

        orig = ho.am_defaults
        ho.am_defaults = lambda v, t0: (cov0.copy(), adapt_when, reset)
        try:
            chain_py, acc = ho.adaptive_metropolis(target, theta0, iters, thin, "hier", rng="philox", seed=seed,
                                                   chain_id=cid)
        finally:
            ho.am_defaults = orig
        lt0 = c_oracle.hier_log_target_batch(ex, theta0[None, :], shapes, scales, locs)[0]
        assert lt0 == pytest.approx(target(theta0), rel=1e-12)
        st = c_oracle.make_state(theta0, lt0, 0.0, cov0)
        chain_c = c_oracle.am_hier(ex, shapes, scales, locs, st, 0, iters, thin, adapt_when, seed, cid)
        assert np.allclose(chain_c, chain_py[1:], rtol=1e-8, atol=1e-8), (drug, channel)
        assert st[-1] / iters == pytest.approx(acc, abs=1e-12)
        assert 0.02 < acc < 0.9


def test_am_zero_start_component_freezes_that_coordinate(table):
    """theta0 with a component of exactly 0 (Sigma0 = 0.05 diag|theta0| then has a zero variance, PyHillFit.py:751):
    the reference's SVD-based draw leaves that coordinate where it is and moves the others.  The Cholesky-based Philox
    path (C oracle == numpy restatement == GPU kernels) stores the zero diagonal as COV0_DIAG_FLOOR: same behaviour,
    no NaN."""
    concs, y = table.concat("Amiodarone", "hERG")
    w0, w100, wo = ho.masks(y)
    pb = ho.compute_pi_bit_of_log_likelihood(wo)
    theta0 = np.array([6.0, 0.0, 7.0])          # Hill exactly 0: in support, zero proposal variance

    def target(th):
        with np.errstate(all="ignore"):
            return ho.log_target(2, y, w0, w100, wo, concs, th, 1, pb)

    chain_py, acc = ho.adaptive_metropolis(target, theta0, 400, 5, "fit", rng="philox", seed=4, chain_id=1)
    assert np.all(np.isfinite(chain_py)) and acc > 0.05
    assert np.all(np.abs(chain_py[:, 1]) < 1e-20)                     # frozen
    assert np.ptp(chain_py[:, 0]) > 1e-3 and np.ptp(chain_py[:, 2]) > 1e-3   # the others move
    cov0, adapt_when, reset = ho.am_defaults("fit", theta0)
    lt0, ll10 = c_oracle.log_target_batch(2, concs, y, theta0[None, :], 1.0, pb)
    st = c_oracle.make_state(theta0, lt0[0], ll10[0], cov0)
    chain_c = c_oracle.am_single(2, concs, y, 1.0, pb, st, 0, 400, 5, adapt_when, reset, 4, 1)
    assert np.allclose(chain_c, chain_py[1:], rtol=1e-9, atol=1e-9)
    # the reference's own draw (numpy multivariate_normal, SVD factor) freezes the coordinate exactly
    import numpy.random as npr
    npr.seed(25)
    chain_np, _ = ho.adaptive_metropolis(target, theta0, 400, 5, "fit", rng="numpy")
    assert np.all(np.abs(chain_np[:, 1]) < 1e-12) and np.ptp(chain_np[:, 0]) > 1e-3


@pytest.mark.parametrize("variant,model,depth", [("temp", 2, 4), ("fit", 2, 2), ("fit", 1, 8), ("temp", 1, 3)])
def test_speculative_round_protocol_commits_the_sequential_chain(table, variant, model, depth):
    """The round protocol of csrc/phf_single_spec.cu, stated in numpy (hill_oracle.adaptive_metropolis_speculative):
    S proposals per round under the hypothesis that their predecessors are rejected, the first accepted one ends the
    round.  It reproduces the sequential loop BIT FOR BIT (chain rows incl. the thinned ones that fall on rejected
    iterations, acceptance), across the start of adaptation and PyHillTemp's mean reset, and needs ~(1 - 0.75^S)/0.25
    times fewer rounds than iterations once the acceptance rate has settled near 0.25."""
    concs, y = table.concat("Amiodarone", "hERG")
    w0, w100, wo = ho.masks(y)
    pb = ho.compute_pi_bit_of_log_likelihood(wo)
    theta0 = np.ones(3 if model == 2 else 2) if variant == "temp" else (np.array([6.0, 0.9, 7.0]) if model == 2
                                                                         else np.array([6.0, 7.0]))
    temp = 0.421875 if variant == "temp" else 1

    def target(th):
        with np.errstate(all="ignore"):
            return ho.log_target(model, y, w0, w100, wo, concs, th, temp, pb)

    cov0, _, reset = ho.am_defaults(variant, theta0)
    orig = ho.am_defaults
    ho.am_defaults = lambda v, t0: (cov0.copy(), 60, reset)      # adaptation (and the mean reset) inside the run
    try:
        iters, thin = 613, 5                                       # not a multiple of the depth or the thinning
        want, acc = ho.adaptive_metropolis(target, theta0, iters, thin, variant, rng="philox", seed=9, chain_id=3)
        got, acc_s, rounds = ho.adaptive_metropolis_speculative(target, theta0, iters, thin, variant, depth, seed=9,
                                                                chain_id=3)
    finally:
        ho.am_defaults = orig
    assert np.array_equal(got, want)
    assert acc_s == pytest.approx(acc, abs=1e-12)   # (a count / n here, a running mean in the sequential loop)
    a = max(acc, 1e-3)
    expect = iters * a / (1 - (1 - a) ** depth)                    # rounds if every iteration accepted with probability a
    assert iters / depth <= rounds <= min(iters, 1.35 * expect + 10)
