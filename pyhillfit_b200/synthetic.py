"""Synthetic dose-response datasets shaped like one drug group of the reference's data/synthetic_data.csv
(5 experiments x doses {0.0008, 0.08, 0.8, 8} uM => N = 20 points, 4 unique doses): BASELINE config 5.

The reference ships the CSV but not its generator; SURVEY.md section 8d fixes this one: pIC50 ~ U[4,8],
Hill ~ U[0.5,2], sigma ~ U[2,10], y = clip(round(curve + sigma*eps, 1), 0, 100) -- one decimal place like the CSV
(data/synthetic_data.csv:2-12), which also produces exact-zero (left-censored) responses like the real file.
"""
import numpy as np

DOSES = np.array([0.0008, 0.08, 0.8, 8.0])
N_EXPTS = 5
SEED = 20161018


def design():
    """concs[20]: experiment-major, as dr.load_crumb_data + the concatenation at PyHillFit.py:661-665 give it."""
    return np.tile(DOSES, N_EXPTS)


def generate(n_datasets, seed=SEED, offset=0):
    """-> (concs[20], responses[n_datasets, 20], truth[n_datasets, 3] = (pIC50, Hill, sigma)).
    Dataset k of a call with `offset` o equals dataset o+k of a call with offset 0 (ranks generate their shard)."""
    concs = design()
    out_y = np.empty((n_datasets, len(concs)))
    truth = np.empty((n_datasets, 3))
    block = 1 << 16
    first = offset // block
    pos = 0
    b = first
    while pos < n_datasets:
        r = np.random.default_rng([seed, b])              # one independent stream per block of 65536 datasets
        pic50 = r.uniform(4, 8, block)
        hill = r.uniform(0.5, 2, block)
        sigma = r.uniform(2, 10, block)
        eps = r.standard_normal((block, len(concs)))
        curve = 100. * (1. - 1. / (1. + (concs[None, :] / 10 ** (6 - pic50[:, None])) ** hill[:, None]))
        y = np.clip(np.round(curve + sigma[:, None] * eps, 1), 0, 100)
        lo = offset - b * block if b == first else 0
        take = min(block - lo, n_datasets - pos)
        out_y[pos:pos + take] = y[lo:lo + take]
        truth[pos:pos + take] = np.stack([pic50, hill, sigma], 1)[lo:lo + take]
        pos += take
        b += 1
    return concs, out_y, truth
