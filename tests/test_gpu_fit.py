"""phf_best_fit_batch (least-squares start points on the device) against the host statement of
the same algorithm (pyhillfit_b200/initial_fit.py: best_fit_batch, itself tied to scalar scipy Nelder-Mead in
tests/test_host_logic.py).  The two evaluate pow() in different libraries, so paths may part by an ulp and the
comparison is on the minimum reached: sum of squares to 1e-6 relative, parameters wherever the minimum is not flat."""
import numpy as np
import pytest

from _data import Table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table():
    return Table("crumb_data")


def _compare(model, data, th, ss, hs_th, hs_ss):
    flat = 0
    for k, (c, y) in enumerate(data):
        assert abs(ss[k] - hs_ss[k]) <= 1e-6 * max(hs_ss[k], 1.0), (k, ss[k], hs_ss[k])
        if np.max(np.abs(th[k] - hs_th[k])) > 1e-4:
            flat += 1
        assert th[k, -1] == pytest.approx(max(np.sqrt(ss[k] / len(c)), 2e-3), rel=1e-12)
        assert th[k, 0] >= -3.0 - 1e-12 and (model == 1 or 0 <= th[k, 1] <= 10.0)
    return flat


@pytest.mark.parametrize("model", [1, 2])
def test_all_crumb_pairs(table, model):
    from pyhillfit_b200.initial_fit import best_fit_batch, best_fit_batch_gpu
    data = [table.concat(*p) for p in table.pairs()]
    assert len(data) == 210
    th, ss = best_fit_batch_gpu(model, data)
    hth, hss = best_fit_batch(model, data)
    flat = _compare(model, data, th, ss, hth, hss)
    assert flat <= len(data) // 3          # the flat ones are the no-block pairs (pIC50 runs to the bound)


def test_per_experiment_fits_of_the_hierarchical_start(table):
    from pyhillfit_b200.initial_fit import best_fit_batch, best_fit_batch_gpu
    ex = [e for p in table.pairs()[::5] for e in table.experiments(*p)]
    data = [(e[:, 0], e[:, 1]) for e in ex]
    th, ss = best_fit_batch_gpu(2, data, pic50_lower=-2.0)
    hth, hss = best_fit_batch(2, data, pic50_lower=-2.0)
    for k in range(len(data)):
        assert abs(ss[k] - hss[k]) <= 1e-6 * max(hss[k], 1.0)
        assert th[k, 0] >= -2.0 - 1e-12


def test_synthetic_config5_datasets_flat_layout_and_truth_recovery():
    """2 000 synthetic config-5 datasets through the (offsets, concs, responses) form: equal to the host fit, and
    the fitted pIC50 sits near the generating one (4 doses x 5 experiments, sigma 2-10 % block)."""
    from pyhillfit_b200 import synthetic
    from pyhillfit_b200.initial_fit import best_fit_batch, best_fit_batch_gpu
    n = 2000
    c20, resp, truth = synthetic.generate(n, seed=11)
    concs = np.tile(c20, (n, 1))
    offsets = np.arange(n + 1, dtype=np.int64) * concs.shape[1]
    th, ss = best_fit_batch_gpu(2, (offsets, concs.reshape(-1), resp.reshape(-1)))
    data = [(concs[k], resp[k]) for k in range(n)]
    hth, hss = best_fit_batch(2, data)
    _compare(2, data, th, ss, hth, hss)
    inside = (truth[:, 0] > 4.5) & (truth[:, 0] < 7.5)
    assert np.median(np.abs(th[inside, 0] - truth[inside, 0])) < 0.15


def test_edge_cases():
    from pyhillfit_b200.initial_fit import best_fit_batch, best_fit_batch_gpu
    assert best_fit_batch_gpu(2, [])[0].shape == (0, 3)
    c = np.array([0.1, 1.0, 10.0, 100.0])
    exact = 100. * (1. - 1. / (1. + (c / 10 ** (6 - 5.5)) ** 1.3))
    data = [(c, exact),                                   # a perfect fit: sigma floors at 2e-3
            (c, np.zeros(4)),                             # no block at all: pIC50 runs to the lower bound
            (c, np.full(4, 100.0)),                       # full block: pIC50 runs up
            (np.array([3.0]), np.array([40.0])),          # one point
            (c, np.array([-12.0, 30.0, 140.0, 95.0]))]    # out-of-range responses count as they are
    th, ss = best_fit_batch_gpu(2, data)
    hth, hss = best_fit_batch(2, data)
    for k in range(len(data)):
        assert abs(ss[k] - hss[k]) <= 1e-6 * max(hss[k], 1.0) + 1e-9
    assert th[0, 0] == pytest.approx(5.5, abs=1e-5) and th[0, 1] == pytest.approx(1.3, abs=1e-5) and th[0, 2] == 2e-3
    assert np.isfinite(th).all()
    with pytest.raises(ValueError):
        best_fit_batch_gpu(2, (np.array([0, 2, 2]), np.ones(2), np.ones(2)))
