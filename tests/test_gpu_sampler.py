"""GPU tests of the fused adaptive-Metropolis kernels: step-by-step trajectory parity with the C oracle on
the shared Philox stream, launch-segmentation invariance, the host-buffer entry point, and statistical
parity with chains produced by the unmodified reference (tests/golden/ref_chains.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

import c_oracle
import hill_oracle as ho
from _data import GOLD, Table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table():
    return Table("crumb_data")


def _oracle_chain(model, concs, y, temp, theta0, cov0, iters, thin, adapt_when, reset, seed, chain_id):
    pb = ho.compute_pi_bit_of_log_likelihood(y)
    lt0, l10 = c_oracle.log_target_batch(model, concs, y, theta0[None, :], temp, pb)
    st = c_oracle.make_state(theta0, lt0[0], l10[0], cov0)
    chain = c_oracle.am_single(model, concs, y, temp, pb, st, 0, iters, thin, adapt_when, reset, seed, chain_id)
    return chain, st


@pytest.mark.parametrize("model,variant", [(2, "temp"), (2, "fit"), (1, "fit"), (1, "temp")])
def test_single_level_trajectories_follow_oracle(table, model, variant):
    """Same seed, same Philox counters, same algorithm: the GPU chain and the C oracle chain agree row by row
    (until fp64 rounding differences flip an accept, which the tolerance window below does not reach)."""
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler, variant_defaults
    pairs = [("Amiodarone", "hERG"), ("Bepridil", "hERG"), ("Amitriptyline", "Kv4.3"), ("Dofetilide", "hERG")]
    pack = SinglePack([table.concat(d, c) for d, c in pairs])
    d = 2 if model == 1 else 3
    temps = [1.0, 0.421875, 0.0, 0.015625] if variant == "temp" else [1.0] * 4
    n_per = 3
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), n_per)
    tt = np.repeat(temps, n_per)
    rng = np.random.default_rng(5)
    if variant == "temp":
        theta0 = np.ones((len(ids), d))
    else:
        theta0 = np.stack([rng.uniform(4.5, 6.5, len(ids)), rng.uniform(0.6, 1.4, len(ids)),
                           rng.uniform(4, 9, len(ids))], 1)
        theta0 = theta0 if model == 2 else theta0[:, [0, 2]]
    iters, thin, adapt_when, seed, base = 600, 5, 100, 77, 1000
    s = SingleLevelSampler(model, pack, ids, tt, theta0, variant=variant, adapt_when=adapt_when, seed=seed,
                           chain_id_base=base, thinning=thin, burn_rows=20)
    row0 = s.initial_row().cpu().numpy()
    got = s.run(iters).cpu().numpy()
    f = s.state_fields()
    cov0, _, reset = variant_defaults(variant, theta0)
    for k in range(len(ids)):
        concs, y = table.concat(*pairs[ids[k]])
        want, st = _oracle_chain(model, concs, y, tt[k], theta0[k], cov0[k], iters, thin, adapt_when, reset, seed,
                                 base + k)
        assert np.allclose(row0[k, :d], theta0[k]) and row0[k, d] == pytest.approx(
            c_oracle.log_target_batch(model, concs, y, theta0[k][None], tt[k],
                                      ho.compute_pi_bit_of_log_likelihood(y))[0][0], rel=1e-12)
        assert np.allclose(got[k], want, rtol=1e-8, atol=1e-8), "chain %d diverged from the oracle" % k
        assert f["n_accepted"][k] == st[-1]
        assert f["loga"][k] == pytest.approx(st[-3], rel=1e-9, abs=1e-9)
        assert np.allclose(f["cov"][k].reshape(-1), st[2 * d + 2:2 * d + 2 + d * d], rtol=1e-7, atol=1e-12)
        assert np.allclose(f["mean"][k], st[d + 2:2 * d + 2], rtol=1e-8)
    # thermodynamic-integration accumulator == mean of the oracle's temperature-1 log-likelihood over rows >= burn
    concs, y = table.concat(*pairs[0])
    want, _ = _oracle_chain(model, concs, y, tt[0], theta0[0], cov0[0], iters, thin, adapt_when, reset, seed, base)
    _, l1 = c_oracle.log_target_batch(model, concs, y, np.ascontiguousarray(want[19:, :d]), 1.0,
                                      ho.compute_pi_bit_of_log_likelihood(y))
    assert s.loglik_t1_mean()[0] == pytest.approx(l1.mean(), rel=1e-9)


def test_segmented_launches_are_bit_identical(table):
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pack = SinglePack([table.concat(d, c) for d, c in table.pairs()[:40]])
    ids = np.repeat(np.arange(40, dtype=np.int32), 5)
    theta0 = np.tile([5.5, 1.0, 6.0], (len(ids), 1))
    kw = dict(variant="fit", adapt_when=300, seed=3, thinning=5, burn_rows=0)
    a = SingleLevelSampler(2, pack, ids, 1.0, theta0, **kw)
    whole = a.run(2000).cpu().numpy()
    b = SingleLevelSampler(2, pack, ids, 1.0, theta0, stage=False, block_threads=64, **kw)
    parts = [b.run(k).cpu().numpy() for k in (5, 700, 33, 1262)]   # 33: segment boundary off the thinning grid
    assert np.array_equal(whole, np.concatenate(parts, axis=1))
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    assert whole.shape == (200, 400, 4)
    acc = a.acceptance()
    assert 0.05 < acc.mean() < 0.6


def test_host_buffer_entry_point(table):
    """phf_am_single_run_host: numpy in, numpy out, identical to the device-pointer path."""
    import torch
    from pyhillfit_b200 import _lib
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pack = SinglePack([table.concat(d, c) for d, c in table.pairs()[:16]])
    ids = np.repeat(np.arange(16, dtype=np.int32), 8)
    theta0 = np.tile([5.5, 1.0, 6.0], (len(ids), 1))
    kw = dict(variant="fit", adapt_when=200, seed=11, thinning=5, burn_rows=10)
    ref = SingleLevelSampler(2, pack, ids, 1.0, theta0, **kw)
    state0 = ref.state.cpu().numpy().copy()
    want = ref.run(1000).cpu().numpy()
    state = torch.from_numpy(state0.copy()).pin_memory().numpy()
    rows = 200
    samples = torch.empty((len(ids), rows, 4), dtype=torch.float64).pin_memory().numpy()
    cfg = _lib.AmConfig(model=2, reset_mean_at_adapt=0, t0=0, n_iters=1000, thinning=5, adapt_when=200, burn_rows=10,
                        rows_capacity=rows, seed=11, chain_id_base=0, stage_groups=0, block_threads=0)
    temps = np.ones(len(ids))
    L = _lib.load()
    _lib.check(L.phf_am_single_run_host(C.byref(cfg), len(ids), state.ctypes.data, ids.ctypes.data,
                                        temps.ctypes.data, pack.n_datasets, pack.datasets.ctypes.data,
                                        len(pack.groups), pack.groups.ctypes.data, samples.ctypes.data, 4, 0),
               "phf_am_single_run_host")
    assert np.array_equal(samples, want)
    assert np.array_equal(state, ref.state.cpu().numpy())
    # row-major samples ([row][chain][d+1]: contiguous segment transfers): the same numbers, transposed
    state_r = torch.from_numpy(state0.copy()).pin_memory().numpy()
    samples_r = torch.empty((rows, len(ids), 4), dtype=torch.float64).pin_memory().numpy()
    cfg.sample_layout = _lib.SAMPLES_ROW_MAJOR
    _lib.check(L.phf_am_single_run_host(C.byref(cfg), len(ids), state_r.ctypes.data, ids.ctypes.data,
                                        temps.ctypes.data, pack.n_datasets, pack.datasets.ctypes.data,
                                        len(pack.groups), pack.groups.ctypes.data, samples_r.ctypes.data, 7, 0),
               "phf_am_single_run_host")
    assert np.array_equal(samples_r.transpose(1, 0, 2), want)
    assert np.array_equal(state_r, state)


@pytest.mark.parametrize("model,lanes", [(2, 1), (2, 2), (1, 4)])
def test_row_major_samples_are_the_transpose(table, model, lanes):
    """cfg.sample_layout = PHF_SAMPLES_ROW_MAJOR changes where rows land, nothing else (device-pointer path, segmented)."""
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pack = SinglePack([table.concat(d, c) for d, c in table.pairs()[:5]])
    ids = np.repeat(np.arange(5, dtype=np.int32), 13)   # 65 chains: a ragged last warp
    d = 2 if model == 1 else 3
    theta0 = np.tile([5.5, 1.0, 6.0] if model == 2 else [5.5, 6.0], (len(ids), 1))
    kw = dict(variant="fit", adapt_when=100, seed=3, thinning=4, lanes=lanes)
    a = SingleLevelSampler(model, pack, ids, 1.0, theta0, **kw)
    b = SingleLevelSampler(model, pack, ids, 1.0, theta0, **kw)
    want = a.run(600).cpu().numpy()
    got = np.concatenate([b.run(250, row_major=True).cpu().numpy(), b.run(350, row_major=True).cpu().numpy()])
    assert got.shape == (150, len(ids), d + 1)
    assert np.array_equal(got.transpose(1, 0, 2), want)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())


@pytest.mark.parametrize("lanes", [0, 1, 4])
def test_hier_trajectories_follow_oracle(table, lanes):
    """lanes = 0: the lane-per-parameter kernel (few chains); lanes = 1: the thread-per-chain kernel; lanes = 4: four
    lanes per chain (at most 5 experiments)."""
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors, variant_defaults
    pr, shapes, scales, locs = hier_priors()
    pairs = table.pairs()
    by_ne = {}
    for ip, (d, c) in enumerate(pairs):
        by_ne.setdefault(len(table.experiments(d, c)), []).append(ip)
    assert sorted(by_ne) == [3, 4, 5, 6]
    for ne, idxs in sorted(by_ne.items()):
        if lanes == 4 and ne > 5:
            continue
        use = idxs[:3]
        pack = HierPack([table.experiments(*pairs[i]) for i in use])
        ids = np.repeat(np.arange(len(use), dtype=np.int32), 2 if lanes != 4 else 3)   # 9 chains: a ragged last warp of quads
        dim = 5 + 2 * ne
        rng = np.random.default_rng(ne)
        theta0 = np.concatenate([np.tile([1.0, 4.0, 6.0, 0.3], (len(ids), 1)),
                                 np.tile([5.0, 1.0], (len(ids), ne)) + rng.uniform(-0.3, 0.3, (len(ids), 2 * ne)),
                                 np.full((len(ids), 1), 8.0)], axis=1)
        iters, thin, adapt_when, seed, base = 300, 5, 60, 5, 10 ** 10
        s = HierarchicalSampler(pack, ids, theta0, pr, adapt_when=adapt_when, seed=seed, chain_id_base=base,
                                thinning=thin, lanes=lanes)
        lt0 = s.initial_row().cpu().numpy()[:, dim]
        got = s.run(iters).cpu().numpy()
        cov0, _, _ = variant_defaults("hier", theta0)
        f = s.state_fields()
        for k in range(len(ids)):
            ex = table.experiments(*pairs[use[ids[k]]])
            want0 = c_oracle.hier_log_target_batch(ex, theta0[k][None], shapes, scales, locs)[0]
            assert lt0[k] == pytest.approx(want0, rel=1e-12)
            st = c_oracle.make_state(theta0[k], want0, 0.0, cov0[k])
            want = c_oracle.am_hier(ex, shapes, scales, locs, st, 0, iters, thin, adapt_when, seed, base + k)
            assert np.allclose(got[k], want, rtol=1e-7, atol=1e-7), "ne=%d chain %d diverged" % (ne, k)
            assert f["n_accepted"][k] == st[-1]


@pytest.mark.parametrize("model", [1, 2])
def test_posterior_quantiles_match_reference_chains(table, model):
    """64 GPU chains per temperature vs the reference's do_mcmc chain (PyHillTemp.py:57-125, numpy RNG): quantiles
    within 4 Monte-Carlo standard errors of the quantile estimates (tail-indicator ESS, tests/_stats.py), the mean
    temperature-1 log-likelihood within 4 combined standard errors."""
    from _stats import assert_quantiles_within_mcse
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    g = np.load(os.path.join(GOLD, "ref_chains.npz"))
    temps = g["temps"]
    sel = [40, 30, 20, 10, 0]
    pack = SinglePack([table.concat("Amiodarone", "hERG")])
    d = 2 if model == 1 else 3
    nch = 64
    tt = np.repeat(temps[sel], nch)
    iters, thin = int(g["iters"]), int(g["thin"])   # same length and burn-in as the reference chains
    burn = (iters // thin + 1) // 4
    s = SingleLevelSampler(model, pack, np.zeros(len(tt), dtype=np.int32), tt, np.ones((len(tt), d)), variant="temp",
                           seed=2024, thinning=thin, burn_rows=burn)
    smp = s.run(iters).cpu().numpy()[:, burn - 1:, :]
    ll1 = s.loglik_t1_mean().reshape(len(sel), nch)
    for j, it in enumerate(sel):
        assert_quantiles_within_mcse(smp[j * nch:(j + 1) * nch, :, :d], g["ladder_m%d_q" % model][it],
                                     g["ladder_m%d_ess_q" % model][it], 4.0, "model %d, temperature %d" % (model, it))
        ref_m, ref_se = g["ladder_m%d_ll1_mean" % model][it], g["ladder_m%d_ll1_sd" % model][it] / np.sqrt(
            g["ladder_m%d_ll1_ess" % model][it])
        gpu_se = ll1[j].std(ddof=1) / np.sqrt(nch)
        assert abs(ll1[j].mean() - ref_m) <= 4 * np.hypot(ref_se, gpu_se) + 1e-9, (it, ll1[j].mean(), ref_m, ref_se)


def test_hier_many_experiments_trajectory():
    """dim 105 (50 experiments, data/synthetic_data.csv): the warp-per-chain sampler follows the C oracle step by
    step, across a launch boundary."""
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors, variant_defaults
    pr, shapes, scales, locs = hier_priors()
    syn = Table("synthetic_data")
    ex = [syn.experiments(d, c) for d, c in syn.pairs() if len(syn.experiments(d, c)) == 50][0]
    ne, dim = 50, 105
    rng = np.random.default_rng(1)
    nch = 3
    theta0 = np.concatenate([np.tile([1.0, 4.0, 6.0, 0.3], (nch, 1)),
                             np.tile([6.0, 1.0], (nch, ne)) + rng.uniform(-0.2, 0.2, (nch, 2 * ne)),
                             np.full((nch, 1), 8.0)], axis=1)
    iters, thin, adapt_when, seed, base = 120, 5, 40, 8, 77
    s = HierarchicalSampler(HierPack([ex]), np.zeros(nch, dtype=np.int32), theta0, pr, adapt_when=adapt_when,
                            seed=seed, chain_id_base=base, thinning=thin)
    got = np.concatenate([s.run(70).cpu().numpy(), s.run(50).cpu().numpy()], axis=1)
    cov0, _, _ = variant_defaults("hier", theta0)
    f = s.state_fields()
    for k in range(nch):
        want0 = c_oracle.hier_log_target_batch(ex, theta0[k][None], shapes, scales, locs)[0]
        st = c_oracle.make_state(theta0[k], want0, 0.0, cov0[k])
        want = c_oracle.am_hier(ex, shapes, scales, locs, st, 0, iters, thin, adapt_when, seed, base + k)
        assert np.allclose(got[k], want, rtol=1e-7, atol=1e-7), "chain %d diverged" % k
        assert f["n_accepted"][k] == st[-1]
    assert f["n_accepted"].sum() > 0


@pytest.mark.parametrize("lanes,block,hint", [(1, 32, 2), (1, 128, 4), (2, 64, 0), (2, 128, 4), (4, 32, 0), (4, 96, 6)])
def test_launch_shapes_do_not_change_results(table, lanes, block, hint):
    """CTA size, register-budget variant and shared-memory staging are tuning knobs: for a fixed lane count every
    combination gives the same bits (ragged chain count: the last warp / CTA is partly idle; one dataset with five
    doses exercises the general-loop path, one with two the predicated-off lanes)."""
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pairs = table.pairs()
    five = [p for p in pairs if len(np.unique(table.concat(*p)[0])) == 5][0]
    two = [p for p in pairs if len(np.unique(table.concat(*p)[0])) == 2][0]
    use = [pairs[0], five, two, pairs[7], pairs[33]]
    pack = SinglePack([table.concat(*p) for p in use])
    ids = np.repeat(np.arange(len(use), dtype=np.int32), 11)[:-2]          # 53 chains
    theta0 = np.tile([5.5, 1.0, 6.0], (len(ids), 1))
    kw = dict(variant="fit", adapt_when=50, seed=5, thinning=5, burn_rows=3, lanes=lanes)
    ref = SingleLevelSampler(2, pack, ids, 1.0, theta0, stage=False, block_threads=32, **kw)
    want = ref.run(333).cpu().numpy()
    s = SingleLevelSampler(2, pack, ids, 1.0, theta0, block_threads=block, **kw)
    s.occupancy_hint = hint
    got = s.run(333).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(s.state.cpu().numpy(), ref.state.cpu().numpy())
    # and the five-dose / two-dose chains still follow the oracle
    from pyhillfit_b200.sampler import variant_defaults
    cov0, _, _ = variant_defaults("fit", theta0)
    for k in (11, 22):
        concs, y = table.concat(*use[ids[k]])
        pb = ho.compute_pi_bit_of_log_likelihood(y)
        lt0, l10 = c_oracle.log_target_batch(2, concs, y, theta0[k][None], 1.0, pb)
        st = c_oracle.make_state(theta0[k], lt0[0], l10[0], cov0[k])
        chain = c_oracle.am_single(2, concs, y, 1.0, pb, st, 0, 333, 5, 50, False, 5, k)
        assert np.allclose(got[k], chain, rtol=1e-8, atol=1e-8)


def test_empty_dataset_and_prior_only_agree(table):
    """A dataset with no observations: the chain samples the prior (pi_bit = 0, no dose groups) exactly like a
    temperature-0 chain of a real dataset with the same Philox ids."""
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pack = SinglePack([table.concat("Amiodarone", "hERG"), (np.zeros(0), np.zeros(0))])
    assert list(pack.datasets["n_groups"]) == [4, 0]
    theta0 = np.ones((4, 3))
    a = SingleLevelSampler(2, pack, np.array([1, 1, 1, 1], dtype=np.int32), 1.0, theta0, variant="temp", seed=2, lanes=4)
    b = SingleLevelSampler(2, pack, np.array([0, 0, 0, 0], dtype=np.int32), 0.0, theta0, variant="temp", seed=2, lanes=4)
    ra, rb = a.run(500).cpu().numpy(), b.run(500).cpu().numpy()
    assert np.allclose(ra, rb, rtol=1e-12, atol=1e-12) and np.all(np.isfinite(ra))


@pytest.mark.parametrize("ne", [3, 4, 6])
def test_hier_thread_kernel_is_resumable_and_matches_the_lane_kernel(table, ne):
    """One thread per chain (cfg.lanes_per_chain = 1): a run cut into two launches is bit-identical to one launch
    (covariance, mean and counters round-trip through the state rows), a ragged last warp and several warps per CTA
    included; and the lane-per-parameter kernel, same seed and stream, gives the same rows to rounding."""
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
    pr, shapes, scales, locs = hier_priors()
    pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne][:3]
    pack = HierPack([table.experiments(*p) for p in pairs])
    per = 2000 if ne == 3 else 50            # 6000 chains: several warps per CTA on every SM
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), per)[:-7]
    rng = np.random.default_rng(ne)
    theta0 = np.concatenate([np.tile([1.0, 4.0, 6.0, 0.3], (len(ids), 1)),
                             np.tile([5.0, 1.0], (len(ids), ne)) + rng.uniform(-0.3, 0.3, (len(ids), 2 * ne)),
                             np.full((len(ids), 1), 8.0)], axis=1)
    kw = dict(adapt_when=40, seed=11, chain_id_base=5, thinning=5)
    a = HierarchicalSampler(pack, ids, theta0, pr, lanes=1, **kw)
    b = HierarchicalSampler(pack, ids, theta0, pr, lanes=1, **kw)
    whole = a.run(100).cpu().numpy()
    parts = np.concatenate([b.run(60).cpu().numpy(), b.run(40).cpu().numpy()], axis=1)
    assert np.array_equal(whole, parts)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    c = HierarchicalSampler(pack, ids[:64], theta0[:64], pr, lanes=16 if ne <= 5 else 32, **kw)
    lane_rows = c.run(100).cpu().numpy()
    same = np.isclose(whole[:64], lane_rows, rtol=1e-9, atol=1e-9).all(axis=(1, 2))
    assert same.mean() >= 0.9, same.mean()    # (a rounding-level difference may flip an accept in a few chains)
    assert 0.05 < a.acceptance().mean() < 0.6


def _hier_setup(table, ne, per, seed=3):
    from pyhillfit_b200.packing import HierPack
    pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne][:3]
    pack = HierPack([table.experiments(*p) for p in pairs])
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), per)[:-3]      # ragged last warp / group
    rng = np.random.default_rng(seed + ne)
    theta0 = np.concatenate([np.tile([1.0, 4.0, 6.0, 0.3], (len(ids), 1)),
                             np.tile([5.0, 1.0], (len(ids), ne)) + rng.uniform(-0.3, 0.3, (len(ids), 2 * ne)),
                             np.full((len(ids), 1), 8.0)], axis=1)
    return pack, ids, theta0


@pytest.mark.parametrize("ne,lanes", [(3, 1), (3, 16), (4, 1), (5, 16), (6, 32)])
def test_hier_row_major_samples_are_the_transpose(table, ne, lanes):
    """cfg.sample_layout = PHF_SAMPLES_ROW_MAJOR in the hierarchical kernels (thread per chain and lane per parameter):
    the same rows, [row][chain][dim+1] instead of [chain][row][dim+1]; state identical; segmented run included."""
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
    pr = hier_priors()[0]
    pack, ids, theta0 = _hier_setup(table, ne, 40)
    kw = dict(adapt_when=40, seed=21, chain_id_base=9, thinning=5, lanes=lanes)
    a = HierarchicalSampler(pack, ids, theta0, pr, **kw)
    b = HierarchicalSampler(pack, ids, theta0, pr, **kw)
    want = a.run(120).cpu().numpy()
    got = np.concatenate([b.run(70, row_major=True).cpu().numpy(), b.run(50, row_major=True).cpu().numpy()], axis=0)
    assert got.shape == (24, len(ids), 5 + 2 * ne + 1)
    assert np.array_equal(got.transpose(1, 0, 2), want)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())


def test_hier_many_experiments_row_major():
    """the warp-per-chain kernel (dim 105) with row-major samples"""
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
    pr = hier_priors()[0]
    syn = Table("synthetic_data")
    ex = [syn.experiments(d, c) for d, c in syn.pairs() if len(syn.experiments(d, c)) == 50][0]
    ne, nch = 50, 3
    rng = np.random.default_rng(1)
    theta0 = np.concatenate([np.tile([1.0, 4.0, 6.0, 0.3], (nch, 1)),
                             np.tile([6.0, 1.0], (nch, ne)) + rng.uniform(-0.2, 0.2, (nch, 2 * ne)),
                             np.full((nch, 1), 8.0)], axis=1)
    kw = dict(adapt_when=40, seed=8, chain_id_base=77, thinning=5)
    a = HierarchicalSampler(HierPack([ex]), np.zeros(nch, dtype=np.int32), theta0, pr, **kw)
    b = HierarchicalSampler(HierPack([ex]), np.zeros(nch, dtype=np.int32), theta0, pr, **kw)
    want = a.run(60).cpu().numpy()
    got = b.run(60, row_major=True).cpu().numpy()
    assert np.array_equal(got.transpose(1, 0, 2), want)


@pytest.mark.parametrize("ne,lanes", [(3, 1), (4, 16)])
def test_hier_host_buffer_entry_point(table, ne, lanes):
    """phf_am_hier_run_host: numpy in, numpy out, identical to the device-pointer path, both sample layouts."""
    import torch
    from pyhillfit_b200 import _lib
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
    pr = hier_priors()[0]
    pack, ids, theta0 = _hier_setup(table, ne, 30)
    dim = 5 + 2 * ne
    kw = dict(adapt_when=50, seed=13, chain_id_base=4, thinning=5, lanes=lanes)
    ref = HierarchicalSampler(pack, ids, theta0, pr, **kw)
    state0 = ref.state.cpu().numpy().copy()
    want = ref.run(400).cpu().numpy()
    rows, n = 80, len(ids)
    L = _lib.load()
    for layout, shape, nseg in ((_lib.SAMPLES_CHAIN_MAJOR, (n, rows, dim + 1), 3),
                                (_lib.SAMPLES_ROW_MAJOR, (rows, n, dim + 1), 7)):
        state = torch.from_numpy(state0.copy()).pin_memory().numpy()
        samples = torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
        cfg = _lib.AmConfig(model=0, reset_mean_at_adapt=0, t0=0, n_iters=400, thinning=5, adapt_when=50,
                            burn_rows=0xFFFFFFFF, rows_capacity=rows, seed=13, chain_id_base=4, stage_groups=0,
                            block_threads=0, lanes_per_chain=lanes, sample_layout=layout)
        _lib.check(L.phf_am_hier_run_host(C.byref(cfg), ne, n, state.ctypes.data, ids.ctypes.data, pack.n_datasets,
                                          pack.datasets.ctypes.data, len(pack.points), pack.points.ctypes.data,
                                          C.byref(pr), samples.ctypes.data, nseg, 0), "phf_am_hier_run_host")
        got = samples if layout == _lib.SAMPLES_CHAIN_MAJOR else samples.transpose(1, 0, 2)
        assert np.array_equal(got, want)
        assert np.array_equal(state, ref.state.cpu().numpy())
    # argument errors come back as codes, not crashes
    assert L.phf_am_hier_run_host(C.byref(cfg), 0, n, state.ctypes.data, ids.ctypes.data, pack.n_datasets,
                                  pack.datasets.ctypes.data, len(pack.points), pack.points.ctypes.data, C.byref(pr),
                                  samples.ctypes.data, 1, 0) != 0
    assert b"n_expts" in L.phf_last_error()


@pytest.mark.parametrize("model,lanes", [(2, 2), (1, 1)])
def test_cta_order_does_not_change_results(table, model, lanes):
    """cfg.cta_order: chain blocks run by decreasing cost (default) or in index order -- which CTA runs a block changes
    nothing: samples and states are bit-identical.  All 210 pairs x 24 chains = several CTAs per SM, a ragged tail."""
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pairs = table.pairs()
    pack = SinglePack([table.concat(d, c) for d, c in pairs])
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), 24)[:-5]
    d = 2 if model == 1 else 3
    theta0 = np.tile([5.5, 1.0, 6.0] if model == 2 else [5.5, 6.0], (len(ids), 1))
    kw = dict(variant="fit", adapt_when=150, seed=4, thinning=5, burn_rows=0, lanes=lanes, block_threads=32)
    a = SingleLevelSampler(model, pack, ids, 1.0, theta0, **kw)
    b = SingleLevelSampler(model, pack, ids, 1.0, theta0, **kw)
    b.cta_order = 1
    ra, rb = a.run(400).cpu().numpy(), b.run(400).cpu().numpy()
    assert ra.shape == (len(ids), 80, d + 1) and np.all(np.isfinite(ra))
    assert np.array_equal(ra, rb)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    rr = a.run(200, row_major=True).cpu().numpy()          # the row-major layout goes through the same block index
    assert np.array_equal(rr.transpose(1, 0, 2), b.run(200).cpu().numpy())


@pytest.mark.parametrize("model,variant,lanes,depth", [(2, "temp", 4, 4), (2, "fit", 2, 4), (1, "temp", 4, 2), (1, "fit", 1, 4),
                                                       (2, "temp", 1, 2), (2, "fit", 4, 8), (1, "temp", 2, 8),
                                                       (2, "temp", 2, 2)])
def test_speculative_evaluation_gives_the_same_chain_bit_for_bit(table, model, variant, lanes, depth):
    """cfg.speculation (csrc/phf_single_spec.cu): `depth` groups of `lanes` lanes evaluate the next `depth` proposals at
    once under the hypothesis that their predecessors are rejected.  Every committed number is computed by the same
    expressions from the same inputs as in the sequential loop, so samples, final state, acceptance counts and the
    thermodynamic-integration accumulator are IDENTICAL to speculation = 1 with the same lane count -- across the start
    of adaptation (and PyHillTemp's mean reset), ragged chain counts, launch boundaries that are not multiples of the
    depth or the thinning, discarded burn-in rows and both sample layouts; and the chain follows the C oracle."""
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler, variant_defaults
    pairs = [("Amiodarone", "hERG"), ("Bepridil", "hERG"), ("Amitriptyline", "Kv4.3"),
             ("Dofetilide", "hERG"),      # five unique doses: the general loop behind the prepared records
             ("Rufinamide", "hERG")]      # two unique doses: absent (all-zero) prepared records
    pack = SinglePack([table.concat(d, c) for d, c in pairs])
    d = 2 if model == 1 else 3
    n_per = 7                                     # 35 chains: ragged for every lanes x depth
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), n_per)
    rng = np.random.default_rng(11)
    tt = rng.choice([1.0, 0.421875, 0.0, 0.015625], len(ids)) if variant == "temp" else np.ones(len(ids))
    if variant == "temp":
        theta0 = np.ones((len(ids), d))
    else:
        theta0 = np.stack([rng.uniform(4.5, 6.5, len(ids)), rng.uniform(0.6, 1.4, len(ids)),
                           rng.uniform(4, 9, len(ids))], 1)
        theta0 = theta0 if model == 2 else theta0[:, [0, 2]]
    kw = dict(variant=variant, adapt_when=120, seed=5, chain_id_base=300, thinning=5, burn_rows=30, lanes=lanes)
    segs = (7, 333, 160, 1, 99)                    # 600 iterations; boundaries off the thinning and the depth
    runs = {}
    for spec in (1, depth):
        for layout in (False, True):
            s = SingleLevelSampler(model, pack, ids, tt, theta0, speculation=spec, **kw)
            assert s.speculation == spec and s.lanes == lanes
            parts = []
            for k in segs:
                r = s.run(k, row_major=layout, discard_burn=layout)
                parts.append((r.transpose(0, 1) if layout else r).cpu().numpy())
            runs[(spec, layout)] = (np.concatenate(parts, axis=1), s.state.cpu().numpy(), s.loglik_t1_mean())
    for layout in (False, True):
        a, b = runs[(1, layout)], runs[(depth, layout)]
        assert a[0].shape == b[0].shape == (len(ids), 120 if not layout else 120 - 29, d + 1)
        assert np.array_equal(a[0], b[0]), "samples differ (layout %s)" % layout
        assert np.array_equal(a[1], b[1]), "final state differs"
        assert np.array_equal(a[2], b[2])
    assert np.array_equal(runs[(depth, False)][0][:, 29:, :], runs[(depth, True)][0])
    # and the chain is the oracle's
    cov0, _, reset = variant_defaults(variant, theta0)
    got = runs[(depth, False)][0]
    for k in (0, 8, 20, 34):
        concs, y = table.concat(*pairs[ids[k]])
        want, st = _oracle_chain(model, concs, y, tt[k], theta0[k], cov0[k], 600, 5, 120, reset, 5, 300 + k)
        assert np.allclose(got[k], want, rtol=1e-8, atol=1e-8), "chain %d diverged from the oracle" % k
        assert runs[(depth, False)][1][k, -1] == st[-1]


def test_speculation_is_chosen_for_small_launches_only(table):
    from pyhillfit_b200 import _lib
    L = _lib.load()
    assert L.phf_am_single_speculation(4000000, 1) == 1       # throughput regime: no speculation
    assert L.phf_am_single_speculation(26880, 2) == 1
    assert L.phf_am_single_speculation(500, 4) >= 4           # a handful of chains: deep speculation


def test_host_entry_point_concurrent_calls_discarded_burn_in_and_release(table):
    """Four host threads call phf_am_single_run_host for the SAME model at once (two workspaces per (device, model):
    two run concurrently, two wait), with cfg.discard_burn_rows = 1: every call returns exactly the post-burn rows of
    the device-pointer path; phf_release_workspaces() frees everything and the next call re-creates what it needs."""
    import threading
    import torch
    from pyhillfit_b200 import _lib
    from pyhillfit_b200.packing import SinglePack
    from pyhillfit_b200.sampler import SingleLevelSampler
    pack = SinglePack([table.concat(d, c) for d, c in table.pairs()[:16]])
    ids = np.repeat(np.arange(16, dtype=np.int32), 8)
    theta0 = np.tile([5.5, 1.0, 6.0], (len(ids), 1))
    iters, thin = 1000, 5
    saved = iters // thin + 1
    burn = saved // 4
    kw = dict(variant="fit", adapt_when=200, seed=11, thinning=thin, burn_rows=burn)
    ref = SingleLevelSampler(2, pack, ids, 1.0, theta0, **kw)
    state0 = ref.state.cpu().numpy().copy()
    full = ref.run(iters).cpu().numpy()
    want = full[:, burn - 1:, :]                      # rows burn .. saved-1 (PyHillFit.py:861-864)
    kept = saved - burn
    assert want.shape[1] == kept
    L = _lib.load()
    temps = np.ones(len(ids))
    out, errs = {}, []

    def worker(k, nseg):
        try:
            state = torch.from_numpy(state0.copy()).pin_memory().numpy()
            samples = torch.zeros((kept, len(ids), 4), dtype=torch.float64).pin_memory().numpy()
            cfg = _lib.AmConfig(model=2, reset_mean_at_adapt=0, t0=0, n_iters=iters, thinning=thin, adapt_when=200,
                                burn_rows=burn, discard_burn_rows=1, rows_capacity=kept, seed=11, chain_id_base=0,
                                lanes_per_chain=ref.lanes, speculation=ref.speculation,
                                sample_layout=_lib.SAMPLES_ROW_MAJOR)
            _lib.check(L.phf_am_single_run_host(C.byref(cfg), len(ids), state.ctypes.data, ids.ctypes.data,
                                                temps.ctypes.data, pack.n_datasets, pack.datasets.ctypes.data,
                                                len(pack.groups), pack.groups.ctypes.data, samples.ctypes.data, nseg, 0),
                       "phf_am_single_run_host")
            out[k] = (samples.transpose(1, 0, 2).copy(), state.copy())
        except Exception as e:
            errs.append(e)

    threads = [threading.Thread(target=worker, args=(k, nseg)) for k, nseg in enumerate((1, 3, 8, 40))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for k in range(4):
        assert np.array_equal(out[k][0], want), k
        assert np.array_equal(out[k][1], ref.state.cpu().numpy()), k
    assert L.phf_release_workspaces() == 0
    worker(9, 5)
    assert not errs and np.array_equal(out[9][0], want)
    # a capacity that ignores the discarded rows is accepted, one below the kept rows is refused
    cfg = _lib.AmConfig(model=2, t0=0, n_iters=iters, thinning=thin, adapt_when=200, burn_rows=burn, discard_burn_rows=1,
                        rows_capacity=kept - 1, seed=11)
    st = state0.copy()
    buf = np.zeros((kept, len(ids), 4))
    assert L.phf_am_single_run_host(C.byref(cfg), len(ids), st.ctypes.data, ids.ctypes.data, temps.ctypes.data,
                                    pack.n_datasets, pack.datasets.ctypes.data, len(pack.groups),
                                    pack.groups.ctypes.data, buf.ctypes.data, 4, 0) == -1


@pytest.mark.parametrize("ne", [3, 4, 5])
def test_hier_four_lane_kernel_segments_layouts_and_ragged_counts(table, ne):
    """The four-lanes-per-chain hierarchical kernel (csrc/phf_hier_quad.cu): launches are resumable bit for bit at
    boundaries off the thinning grid, the row-major layout is the transpose, a chain count that fills neither a warp of
    quads nor a CTA changes nothing for the chains that are there, and the library picks this form for mid-size
    launches."""
    from pyhillfit_b200 import _lib
    from pyhillfit_b200.packing import HierPack
    from pyhillfit_b200.sampler import HierarchicalSampler, hier_priors
    pr, shapes, scales, locs = hier_priors()
    pairs = [p for p in table.pairs() if len(table.experiments(*p)) == ne][:4]
    pack = HierPack([table.experiments(*p) for p in pairs])
    ids = np.repeat(np.arange(len(pairs), dtype=np.int32), 11)[:-3]            # 41 chains
    rng = np.random.default_rng(ne)
    theta0 = np.concatenate([np.tile([1.0, 4.0, 6.0, 0.3], (len(ids), 1)),
                             np.tile([5.0, 1.0], (len(ids), ne)) + rng.uniform(-0.3, 0.3, (len(ids), 2 * ne)),
                             np.full((len(ids), 1), 8.0)], axis=1)
    kw = dict(adapt_when=50, seed=5, chain_id_base=77, thinning=5, lanes=4)
    a = HierarchicalSampler(pack, ids, theta0, pr, **kw)
    whole = a.run(400).cpu().numpy()
    b = HierarchicalSampler(pack, ids, theta0, pr, block_threads=64, **kw)
    parts = [b.run(k, row_major=True).cpu().numpy().transpose(1, 0, 2) for k in (7, 131, 262)]
    assert np.array_equal(whole, np.concatenate(parts, axis=1))
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    c = HierarchicalSampler(pack, ids[:17], theta0[:17], pr, **kw)              # the same chains in a smaller launch
    assert np.array_equal(c.run(400).cpu().numpy(), whole[:17])
    assert 0.05 < a.acceptance().mean() < 0.6 and np.all(np.isfinite(whole))
    L = _lib.load()
    sms = 148
    assert L.phf_am_hier_lanes(ne, 8 * sms) in (16, 32)
    assert L.phf_am_hier_lanes(ne, 64 * sms) == 4
    assert L.phf_am_hier_lanes(ne, 300 * sms) == (1 if ne <= 3 else 4)
    assert L.phf_am_hier_lanes(6, 300 * sms) == 32 and L.phf_am_hier_lanes(50, 10) == 32
