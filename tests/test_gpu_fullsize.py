"""BASELINE.json's full sizes through size-independent properties (the oracle is too slow there):
  * every saved row's log-target column equals the batched log-target kernel evaluated on the row's parameters
    (a checksum of the sampler against the separately parity-tested target kernel);
  * launch segmentation and CTA shape do not change a single bit;
  * 1, 2 and 4 lanes per chain agree to rounding;
  * different chains of one dataset differ, the same (seed, chain id) reproduces exactly."""
import numpy as np
import pytest

from _data import Table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table():
    return Table("crumb_data")


@pytest.fixture(scope="module")
def crumb_pack(table):
    from pyhillfit_b200.packing import SinglePack
    return SinglePack([table.concat(d, c) for d, c in table.pairs()])


def _rows_match_target(model, pack, ids, temps, samples, tol=1e-12):
    from pyhillfit_b200.sampler import log_target_batch
    n, rows, w = samples.shape
    th = samples[:, :, :w - 1].reshape(-1, w - 1)
    lt, _ = log_target_batch(model, pack, th, np.repeat(ids, rows), np.repeat(temps, rows))
    lt = lt.cpu().numpy().reshape(n, rows)
    got = samples[:, :, w - 1]
    assert np.all(np.isfinite(got))
    err = np.abs(got - lt) / np.maximum(1.0, np.abs(lt))
    assert err.max() <= tol, err.max()


@pytest.mark.parametrize("model", [1, 2])
def test_config2_all_pairs_64_chains(table, crumb_pack, model):
    """210 pairs x 64 chains (13 440 chains per model), PyHillFit variant."""
    from pyhillfit_b200.sampler import SingleLevelSampler
    d = 2 if model == 1 else 3
    ids = np.repeat(np.arange(210, dtype=np.int32), 64)
    rng = np.random.default_rng(model)
    theta0 = np.stack([rng.uniform(4.5, 6.5, len(ids)), rng.uniform(0.6, 1.4, len(ids)), rng.uniform(4, 9, len(ids))], 1)
    theta0 = theta0 if model == 2 else theta0[:, [0, 2]]
    kw = dict(variant="fit", adapt_when=200, seed=25, thinning=5, burn_rows=0)
    a = SingleLevelSampler(model, crumb_pack, ids, 1.0, theta0, **kw)
    assert a.lanes == 2                       # more than one warp per sub-partition: two lanes (phf_am_single_lanes)
    whole = a.run(1000).cpu().numpy()
    assert whole.shape == (13440, 200, d + 1)
    _rows_match_target(model, crumb_pack, ids, np.ones(len(ids)), whole[:, ::20, :])
    # segmentation + CTA shape invariance, bit for bit
    b = SingleLevelSampler(model, crumb_pack, ids, 1.0, theta0, block_threads=64, stage=False, **kw)
    parts = np.concatenate([b.run(k).cpu().numpy() for k in (5, 333, 662)], axis=1)
    assert np.array_equal(whole, parts) and np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    # chains of one pair are distinct, reruns reproduce
    assert len({whole[k, -1, 0] for k in range(64)}) > 32
    c = SingleLevelSampler(model, crumb_pack, ids, 1.0, theta0, **kw)
    assert np.array_equal(c.run(1000).cpu().numpy(), whole)
    acc = a.acceptance()
    assert 0.05 < acc.mean() < 0.7


def test_lane_counts_agree_to_rounding(table, crumb_pack):
    from pyhillfit_b200.sampler import SingleLevelSampler
    ids = np.repeat(np.arange(210, dtype=np.int32), 4)
    theta0 = np.tile([5.5, 1.0, 6.0], (len(ids), 1))
    runs = {}
    for lanes in (1, 2, 4):
        s = SingleLevelSampler(2, crumb_pack, ids, 1.0, theta0, variant="fit", adapt_when=100, seed=3, thinning=5,
                               lanes=lanes)
        runs[lanes] = s.run(300).cpu().numpy()
    for lanes in (2, 4):
        same = np.isclose(runs[lanes], runs[1], rtol=1e-9, atol=1e-9).all(axis=(1, 2))
        assert same.mean() > 0.995, (lanes, same.mean())   # a rounding-level accept flip may fork a chain; rare


def test_config4_64_temperature_ladder(table, crumb_pack):
    """PyHillTemp variant, 64 temperatures x 210 pairs (13 440 chains per model): prior-only chains (t = 0) never see
    the data, every row's log-target is the tempered target of its parameters."""
    from pyhillfit_b200 import ti
    from pyhillfit_b200.sampler import SingleLevelSampler
    temps = (np.arange(64.) / 63) ** 3
    ids, tt = ti.build_chain_list(210, temps, 1)
    s = SingleLevelSampler(2, crumb_pack, ids, tt, np.ones((len(ids), 3)), variant="temp", seed=1, thinning=5,
                           burn_rows=50)
    smp = s.run(1000).cpu().numpy()
    _rows_match_target(2, crumb_pack, ids, tt, smp[:, ::25, :])
    t0 = tt == 0.0
    prior = -0.2 * smp[t0, :, 0] + 4 * np.log(smp[t0, :, 2] - 1e-3) - (smp[t0, :, 2] - 1e-3) / 1.49975
    assert np.allclose(smp[t0, :, 3], prior, rtol=1e-12, atol=1e-12)
    means = s.loglik_t1_mean().reshape(210, 64)
    assert np.all(np.isfinite(means))
    assert np.median(means[:, -1] - means[:, 0]) > 0      # hotter chains fit the data worse


def test_config3_hierarchical_every_pair(table):
    """Hierarchical model for every Crumb pair (Ne = 3..6, dim 11..17), 256 chains each for one group of pairs and
    8 for the rest: the log-target column equals the hierarchical target kernel on the row."""
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_log_target_batch, hier_priors
    pr, shapes, scales, locs = hier_priors()
    pairs = table.pairs()
    by_ne = {}
    for ip, (d, c) in enumerate(pairs):
        by_ne.setdefault(len(table.experiments(d, c)), []).append(ip)
    total = 0
    for ne, idxs in sorted(by_ne.items()):
        per = 256 if ne == 6 else 8
        pack = HierPack([table.experiments(*pairs[i]) for i in idxs])
        ids = np.repeat(np.arange(len(idxs), dtype=np.int32), per)
        dim = 5 + 2 * ne
        theta0 = np.tile(np.concatenate(([1.0, 4.0, 6.0, 0.3], np.tile([5.5, 1.0], ne), [8.0])), (len(ids), 1))
        s = HierarchicalSampler(pack, ids, theta0, pr, seed=ne, thinning=5, adapt_when=100)
        smp = s.run(400).cpu().numpy()
        total += len(ids)
        rows = smp[:, ::16, :]
        th = np.ones((rows.shape[0] * rows.shape[1], 17))
        th[:, :dim] = rows[:, :, :dim].reshape(-1, dim)
        want = hier_log_target_batch(pack, th, np.repeat(ids, rows.shape[1]), pr).cpu().numpy()
        got = rows[:, :, dim].reshape(-1)
        assert np.all(np.isfinite(got))
        assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) <= 1e-12
        assert 0.02 < s.acceptance().mean() < 0.8
    assert total == 154 * 8 + 41 * 8 + 12 * 8 + 3 * 256


def test_config5_synthetic_scale_up_shard():
    """One GPU's slice of the 1M-dataset synthetic sweep at 1/8 scale: 62 500 datasets x 4 chains, model 2,
    one thread per chain."""
    from pyhillfit_b200 import synthetic
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    concs, Y, truth = synthetic.generate(62500, offset=125000)
    pack = SinglePack.from_uniform(concs, Y)
    ids = np.repeat(np.arange(62500, dtype=np.int32), 4)
    theta0 = np.tile([6.0, 1.0, 6.0], (len(ids), 1))
    s = SingleLevelSampler(2, pack, ids, 1.0, theta0, variant="fit", seed=9, thinning=5, adapt_when=300)
    assert s.lanes == 1
    s.run(1500, keep=False)
    smp = s.run(500).cpu().numpy()
    assert smp.shape == (250000, 100, 4)
    pick = np.arange(0, 250000, 997)
    _rows_match_target(2, pack, ids[pick], np.ones(len(pick)), smp[pick][:, ::10, :])
    # posterior means recover the generating parameters on average (pIC50 within the prior's pull)
    est = smp[:, :, 0].mean(axis=1).reshape(62500, 4).mean(axis=1)
    assert abs(np.median(est - truth[:, 0])) < 0.1
