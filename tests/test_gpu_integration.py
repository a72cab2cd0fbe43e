"""The binding documented in INTEGRATION.md (examples/reference_binding.py: plain ctypes + numpy, nothing from the
package) gives the same numbers as the package's own host mirror."""
import importlib.util
import os

import numpy as np
import pytest

from _data import Table

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub():
    spec = importlib.util.spec_from_file_location("reference_binding", os.path.join(ROOT, "examples", "reference_binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_documented_binding_matches_the_host_mirror():
    import pyhillfit_b200.doseresponse as dr
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    rb = _stub()
    t = Table("crumb_data")
    concs, y = t.concat("Amiodarone", "hERG")
    w0, w100, wo = y == 0, y == 100, (0 < y) & (y < 100)
    pb = dr.compute_pi_bit_of_log_likelihood(wo)
    dr.define_model(2)
    th = np.array([5.5, 0.8, 8.0])
    got = rb.log_target(2, y, w0, w100, wo, concs, th, 0.3, pb)
    assert got == dr.log_target(y, w0, w100, wo, concs, th, 0.3, pb)
    assert rb.log_target(2, y, w0, w100, wo, concs, np.array([6, 1, 5.]), 1, pb) == pytest.approx(-58.39140921642633,
                                                                                                 rel=1e-12)
    # PyHillTemp-style run: 5 temperatures, theta0 = ones, cov0 = I, mean reset at adapt_when
    temps = (np.arange(41.) / 40) ** 3
    sel = temps[[0, 10, 20, 30, 40]]
    theta0 = np.ones(3)
    pack = SinglePack([(concs, y)])
    ref = SingleLevelSampler(2, pack, np.zeros(5, dtype=np.int32), sel, np.ones((5, 3)), variant="temp", seed=1,
                             adapt_when=300, thinning=5, lanes=4)
    row0 = ref.initial_row().cpu().numpy()
    want = ref.run(2000).cpu().numpy()
    lt0 = row0[:, 3]
    ll1 = ref.state_fields()["loglik_t1"] * 0 + np.array([rb.log_target(2, y, w0, w100, wo, concs, theta0, 1, pb)] * 5)
    chain, state = rb.run_single_level_loop(2, concs, y, theta0, np.eye(3), lt0, ll1, sel, 2000, 5, 300, variant="temp",
                                            seed=1)
    assert chain.shape == (5, 401, 4)
    assert np.array_equal(chain[:, 0, :], row0)
    assert np.allclose(chain[:, 1:, :], want, rtol=1e-9, atol=1e-9)     # the host call picks its own lane count


def test_documented_best_fit_binding_matches_the_package():
    from pyhillfit_b200.initial_fit import best_fit_batch_gpu
    rb = _stub()
    t = Table("crumb_data")
    data = [t.concat(*p) for p in t.pairs()[:12]]
    for model in (1, 2):
        th, ss = rb.best_fits(model, data)
        wt, ws = best_fit_batch_gpu(model, data)
        assert np.array_equal(th, wt) and np.array_equal(ss, ws)
