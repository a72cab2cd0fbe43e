"""On-disk chain / sample files, byte-compatible with what the reference writes and its consumers read.

Formats (every consumer uses np.loadtxt, '#' lines are comments):
  single-level, PyHillFit : 1 header line, rows (pIC50,[Hill,]sigma,log-target), burn-in removed
                            (python/PyHillFit.py:861-867)
  single-level, PyHillTemp: no header, burn-in removed (python/PyHillTemp.py:125,169)
  hierarchical            : 2 header lines, whole chain incl. burn-in (python/PyHillFit.py:424-426,514-515)
  (alpha,mu) samples      : 1 header line, num_APs random post-burn rows, columns [0,2] (python/PyHillFit.py:519-525)
  best_fit_params.txt     : python/PyHillFit.py:739-746
  BFs/<drug>_<channel>_B12.txt : python/compute_bayes_factors.py:86-100
All numbers are written with numpy's default '%.18e', space separated.
"""
import os

import numpy as np

from . import _lib


def save_single_level_chain(chain_file, chain, drug, channel):
    # (the reference's header names the columns in the wrong order; kept verbatim)
    _lib.write_rows_text(chain_file, chain, header='# Nonhierarchical MCMC output for {} + {}: '
                         '(Hill,pIC50,sigma,log-target)\n'.format(drug, channel))


def save_tempered_chain(chain_file, chain):
    _lib.write_rows_text(chain_file, chain)


def save_hierarchical_chain(chain_file, chain):
    _lib.write_rows_text(chain_file, chain,
                         header="# Hill ~ log-logistic(alpha,beta), pIC50 ~ logistic(mu,s)\n"
                                "# alpha, beta, mu, s, hill_1, pic50_1, hill_2, pic50_2, ..., hill_Ne, pic50_Ne, sigma\n")


def save_alpha_mu_samples(samples_file, chain, burn, num_APs, drug, channel, rng=None):
    rng = np.random if rng is None else rng
    saved_iterations = chain.shape[0]
    indices = rng.randint(burn, saved_iterations, num_APs)
    with open(samples_file, 'w') as outfile:
        outfile.write('# {} (alpha,mu) samples from hierarchical MCMC for {} + {}\n'.format(num_APs, drug, channel))
        np.savetxt(outfile, chain[indices, :][:, [0, 2]])


def save_best_fit_params(images_dir, drug, channel, model, theta):
    best_params_file = images_dir + "{}_{}_best_fit_params.txt".format(drug, channel)
    with open(best_params_file, "w") as outfile:
        outfile.write("# CMA-ES best fit params\n")
        if model == 1:
            outfile.write("# pIC50, sigma, (Hill=1, not included)\n")
        elif model == 2:
            outfile.write("# pIC50, Hill, sigma\n")
        np.savetxt(outfile, [theta])
    return best_params_file


def save_bayes_factor(drug, channel, Bij, bf_dir="BFs/"):
    if not os.path.exists(bf_dir):
        os.makedirs(bf_dir)
    bf_file = bf_dir + "{}_{}_B12.txt".format(drug, channel)
    np.savetxt(bf_file, [Bij])
    return bf_file


def extra_chain_name(chain_file, k):
    """File for replicate chain k > 0 of the same target (the reference runs one chain; chain 0 keeps its name)."""
    root, ext = os.path.splitext(chain_file)
    return "{}_rep{}{}".format(root, k, ext)
