"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libphf_b200.so; the oracle is only the checker."""
import os

import numpy as np
import pytest

import c_oracle
import hill_oracle as ho
from _data import GOLD, Table

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # north-star: |gpu - ref| <= 1e-12 * max(1, |ref|), +-inf exact


def assert_close(a, b, tol=RTOL, what=""):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape
    inf = ~np.isfinite(b)
    assert np.array_equal(a[inf], b[inf]), "%s: non-finite values differ" % what
    err = np.abs(a[~inf] - b[~inf]) / np.maximum(1.0, np.abs(b[~inf]))
    assert err.size == 0 or err.max() <= tol, "%s: max scaled error %.3e at %d" % (what, err.max(), err.argmax())


@pytest.fixture(scope="module")
def table():
    return Table("crumb_data")


@pytest.fixture(scope="module")
def single_pack(table):
    from pyhillfit_b200.packing import SinglePack
    return SinglePack([table.concat(d, c) for d, c in table.pairs()])


@pytest.fixture(scope="module")
def hier_pack(table):
    from pyhillfit_b200.packing import HierPack
    return HierPack([table.experiments(d, c) for d, c in table.pairs()])


def test_library_loaded_and_fp64_probe():
    from pyhillfit_b200 import _lib
    assert _lib.load().phf_version() == 100
    tf, sec = _lib.fp64_peak_tflops(3)
    assert 5.0 < tf < 80.0, tf
    assert _lib.launch_count() > 0


@pytest.mark.parametrize("model", [1, 2])
def test_log_target_matches_reference_golden(single_pack, model):
    """All 210 Crumb pairs x 48 parameter vectors (incl. out-of-support and overflow cases) x 8 temperatures,
    against outputs of the unmodified reference."""
    from pyhillfit_b200.sampler import log_target_batch
    g = np.load(os.path.join(GOLD, "log_target_golden.npz"))
    th = g["theta_m%d" % model]
    npairs, nt, d = th.shape
    ids = np.repeat(np.arange(npairs, dtype=np.int32), nt)
    lt, l1 = log_target_batch(model, single_pack, th.reshape(-1, d), ids, g["t_m%d" % model].reshape(-1))
    assert_close(lt.cpu().numpy().reshape(npairs, nt), g["log_target_m%d" % model], what="log_target")
    assert_close(l1.cpu().numpy().reshape(npairs, nt), g["log_lik_t1_m%d" % model], what="loglik_t1")


@pytest.mark.parametrize("model", [1, 2])
def test_log_target_matches_oracle_random(table, single_pack, model):
    """>= 1e4 random theta per (pair, model) for 21 pairs, 2e3 for every other pair, vs the C oracle."""
    from pyhillfit_b200.sampler import log_target_batch
    rng = np.random.default_rng(1234 + model)
    d = 2 if model == 1 else 3
    ladder = ho.temperature_ladder()
    for ip, (drug, channel) in enumerate(table.pairs()):
        n = 10000 if ip % 10 == 0 else 2000
        pic50 = rng.uniform(-3.5, 11, n)
        hill = rng.uniform(-0.2, 10.2, n)
        sigma = np.exp(rng.uniform(np.log(8e-4), np.log(50.), n))
        half = n // 2
        pic50[:half] = rng.uniform(3, 9, half)
        hill[:half] = rng.uniform(0.3, 3, half)
        sigma[:half] = rng.uniform(1, 15, half)
        th = np.stack([pic50, hill, sigma], 1) if model == 2 else np.stack([pic50, sigma], 1)
        tt = ladder[rng.integers(0, len(ladder), n)]
        concs, y = table.concat(drug, channel)
        want, want1 = c_oracle.log_target_batch(model, concs, y, th, tt, ho.compute_pi_bit_of_log_likelihood(y))
        lt, l1 = log_target_batch(model, single_pack, th, np.full(n, ip, dtype=np.int32), tt)
        assert_close(lt.cpu().numpy(), want, what="%s/%s" % (drug, channel))
        assert_close(l1.cpu().numpy(), want1, what="%s/%s ll1" % (drug, channel))


@pytest.mark.parametrize("model", [1, 2])
def test_log_target_synthetic_config5_datasets_vs_oracle(model):
    """BASELINE config 5's data path: 12 000 synthetic datasets (5 experiments x 4 doses, responses rounded to one
    decimal and clipped to [0, 100]) packed by the VECTORISED packer SinglePack.from_uniform, two parameter vectors
    per dataset (one near the generating values, one anywhere incl. out of support), random ladder temperatures --
    against the C oracle evaluated on each dataset's raw (concs, responses)."""
    from pyhillfit_b200 import synthetic
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import log_target_batch
    n_ds = 12000
    concs, Y, truth = synthetic.generate(n_ds, offset=777)
    pack = SinglePack.from_uniform(concs, Y)
    rng = np.random.default_rng(55 + model)
    near = truth * (1.0 + 0.05 * rng.standard_normal(truth.shape))
    far = np.stack([rng.uniform(-3.5, 11, n_ds), rng.uniform(-0.2, 10.2, n_ds),
                    np.exp(rng.uniform(np.log(8e-4), np.log(50.), n_ds))], 1)
    th = np.concatenate([near, far])
    th = th if model == 2 else np.ascontiguousarray(th[:, [0, 2]])
    ids = np.tile(np.arange(n_ds, dtype=np.int32), 2)
    ladder = ho.temperature_ladder()
    tt = ladder[rng.integers(0, len(ladder), len(ids))]
    tt[:n_ds:3] = 1.0
    lt, l1 = log_target_batch(model, pack, th, ids, tt)
    lt, l1 = lt.cpu().numpy(), l1.cpu().numpy()
    want, want1 = np.empty(len(ids)), np.empty(len(ids))
    pb = ho.compute_pi_bit_of_log_likelihood(concs)
    for k in range(len(ids)):
        w, w1 = c_oracle.log_target_batch(model, concs, Y[ids[k]], th[k][None], tt[k], pb)
        want[k], want1[k] = w[0], w1[0]
    assert np.isfinite(want[:n_ds]).mean() > 0.95 and (Y == 0).sum() > 1000 and (Y == 100).sum() > 100
    assert_close(lt, want, what="synthetic log_target")
    assert_close(l1, want1, what="synthetic loglik_t1")


def test_doseresponse_scalar_api(table):
    """The reference's own call signatures (dr.log_target & co.) on SURVEY 8c's known answers."""
    import pyhillfit_b200.doseresponse as dr
    concs, y = table.concat("Amiodarone", "hERG")
    w0, w100, wo = y == 0, y == 100, (0 < y) & (y < 100)
    pb = dr.compute_pi_bit_of_log_likelihood(wo)
    dr.define_model(2)
    assert dr.num_params == 3
    for th, t, want in [((6, 1, 5.), 1, -58.39140921642633), ((6, 1, 5.), 0.125, -5.633162956781316),
                        ((6, 1, 5.), 0, 1.9037293660251158), ((5.5, 0.8, 8.), 1, -55.87532519931496),
                        ((1, 1, 1.), 1, -12946.405169647265), ((400, 1, 5.), 1, -1331.5327403576296)]:
        got = dr.log_target(y, w0, w100, wo, concs, np.array(th, dtype=float), t, pb)
        assert got == pytest.approx(want, rel=1e-12)
    assert dr.log_target(y, w0, w100, wo, concs, np.array([-3.5, 11, 5e-4]), 1, pb) == -np.inf
    assert dr.log_target(y, w0, w100, wo, concs, np.array([6, 1, 1e-3]), 1, pb) == -np.inf
    assert dr.log_data_likelihood(y, w0, w100, wo, concs, np.array([6, 1, 5.]), 0, pb) == 0
    th = np.array([5.5, 0.8, 8.])
    with np.errstate(all="ignore"):
        assert dr.log_priors(th) == pytest.approx(ho.log_priors(2, th), rel=1e-13)
        assert dr.log_data_likelihood(y, w0, w100, wo, concs, th, 0.3, pb) == pytest.approx(
            ho.log_data_likelihood(2, y, w0, w100, wo, concs, th, 0.3, pb), rel=1e-12)
    assert dr.log_priors(np.array([5.5, 10.5, 8.])) == -np.inf
    dr.define_model(1)
    assert dr.log_target(y, w0, w100, wo, concs, np.array([5.5, 8.]), 1, pb) == pytest.approx(-63.44284177305924,
                                                                                             rel=1e-12)


def test_hier_log_target_matches_reference_golden(hier_pack):
    from pyhillfit_b200.sampler import hier_log_target_batch, hier_priors
    g = np.load(os.path.join(GOLD, "hier_target_golden.npz"))
    pr, shapes, scales, locs = hier_priors()
    assert np.array_equal(shapes, g["shapes"]) and np.array_equal(scales, g["scales"])
    th = g["theta"]
    npairs, nt, stride = th.shape
    ids = np.repeat(np.arange(npairs, dtype=np.int32), nt)
    flat = np.nan_to_num(th.reshape(-1, stride), nan=1.0)
    got = hier_log_target_batch(hier_pack, flat, ids, pr).cpu().numpy().reshape(npairs, nt)
    assert_close(got, g["log_target"], what="hier log_target")


def test_hier_log_target_matches_oracle_random(table, hier_pack):
    from pyhillfit_b200.sampler import hier_log_target_batch, hier_priors
    pr, shapes, scales, locs = hier_priors()
    rng = np.random.default_rng(99)
    for ip, (drug, channel) in enumerate(table.pairs()):
        ex = table.experiments(drug, channel)
        ne = len(ex)
        dim = 5 + 2 * ne
        n = 4000 if ip % 10 == 0 else 500
        th = np.zeros((n, dim))
        th[:, 0] = rng.uniform(0.05, 3, n)
        th[:, 1] = rng.uniform(1.95, 12, n)
        th[:, 2] = rng.uniform(-4.2, 10, n)
        th[:, 3] = rng.uniform(0.009, 2, n)
        th[:, 4:-1:2] = rng.uniform(-2.05, 10, (n, ne))
        th[:, 5:-1:2] = rng.uniform(-0.02, 5, (n, ne))
        th[:, -1] = np.exp(rng.uniform(np.log(0.05), np.log(40.), n))
        want = c_oracle.hier_log_target_batch(ex, th, shapes, scales, locs)
        pad = np.ones((n, 17))
        pad[:, :dim] = th
        got = hier_log_target_batch(hier_pack, pad, np.full(n, ip, dtype=np.int32), pr).cpu().numpy()
        assert_close(got, want, what="%s/%s" % (drug, channel))
        assert np.isfinite(want).mean() > 0.2


def test_hier_many_experiments_target(table):
    """The warp-per-chain path (dim > 31): the 50-experiment group of data/synthetic_data.csv (dim 105) against the C
    oracle, and the same kernel on Crumb pairs (theta rows padded beyond 31) against the reference golden values."""
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import hier_log_target_batch, hier_priors
    pr, shapes, scales, locs = hier_priors()
    syn = Table("synthetic_data")
    by_ne = {len(syn.experiments(d, c)): (d, c) for d, c in syn.pairs()}
    assert 50 in by_ne
    ex = syn.experiments(*by_ne[50])
    ne, dim = 50, 105
    rng = np.random.default_rng(50)
    n = 600
    th = np.zeros((n, dim))
    th[:, 0] = rng.uniform(0.05, 3, n)
    th[:, 1] = rng.uniform(1.95, 12, n)
    th[:, 2] = rng.uniform(-4.2, 10, n)
    th[:, 3] = rng.uniform(0.009, 2, n)
    th[:, 4:-1:2] = rng.uniform(3, 9, (n, ne))
    th[:, 5:-1:2] = rng.uniform(0.2, 3, (n, ne))
    th[:30, 4] = rng.uniform(-2.1, -1.9, 30)           # around the pIC50 support edge
    th[30:60, 7] = rng.uniform(-0.01, 0.01, 30)        # around the Hill support edge
    th[:, -1] = np.exp(rng.uniform(np.log(0.05), np.log(40.), n))
    want = c_oracle.hier_log_target_batch(ex, th, shapes, scales, locs)
    pack = HierPack([ex])
    got = hier_log_target_batch(pack, th, np.zeros(n, dtype=np.int32), pr).cpu().numpy()
    assert_close(got, want, what="synthetic 50 experiments")
    assert 0.3 < np.isfinite(want).mean() < 1.0
    # same kernel, small dims: golden values of the unmodified reference
    g = np.load(os.path.join(GOLD, "hier_target_golden.npz"))
    cpack = HierPack([table.experiments(d, c) for d, c in table.pairs()])
    npairs, nt, stride = g["theta"].shape
    pad = np.ones((npairs * nt, 40))
    pad[:, :stride] = np.nan_to_num(g["theta"].reshape(-1, stride), nan=1.0)
    got = hier_log_target_batch(cpack, pad, np.repeat(np.arange(npairs, dtype=np.int32), nt), pr).cpu().numpy()
    assert_close(got.reshape(npairs, nt), g["log_target"], what="warp-per-vector kernel on Crumb pairs")


def test_posterior_predictive_cdfs_match_reference():
    """phf_hier_predictive_cdfs vs the unmodified reference function (tests/golden/cdf_golden.npz: 3000 rows incl. a
    logistic narrower than the grid step and a very steep log-logistic) and vs the oracle on a strided chain buffer.
    Tolerance 1e-11 relative (+1e-300): the reference sums 3000 terms sequentially, the kernel in 128-row tiles."""
    import torch
    from pyhillfit_b200.construct_hierarchical_cdfs import (construct_posterior_predictive_cdfs,
                                                            predictive_cdfs_from_rows)
    g = np.load(os.path.join(GOLD, "cdf_golden.npz"))
    r = g["rows"]
    hx, hc, px, pc, hp, pp = construct_posterior_predictive_cdfs(r[:, 0], r[:, 1], r[:, 2], r[:, 3])
    assert np.array_equal(hx, g["hill_x"]) and np.array_equal(px, g["pic50_x"])
    for got, key in ((hc, "hill_cdf"), (pc, "pic50_cdf"), (hp, "hill_pdf"), (pp, "pic50_pdf")):
        want = g[key]
        assert np.all(np.abs(got - want) <= 1e-11 * np.abs(want) + 1e-300), (key, np.max(np.abs(got - want)))
    assert hc[0] == 0.0 and abs(pc[-1] - 1) < 0.01
    # a chain-shaped device buffer (12 columns, first four used), few rows
    rng = np.random.default_rng(3)
    chain = rng.uniform(0.5, 3.0, (37, 12))
    chain[:, 1] += 2.0
    with np.errstate(all="ignore"):
        want = ho.construct_posterior_predictive_cdfs(chain[:, 0], chain[:, 1], chain[:, 2], chain[:, 3])
    got = predictive_cdfs_from_rows(torch.from_numpy(chain).cuda())
    for k, j in ((0, 1), (1, 4), (2, 3), (3, 5)):
        assert np.allclose(got[k], want[j], rtol=1e-12, atol=1e-300)
