"""compute_bayes_factors command line (python/compute_bayes_factors.py of the reference).

Reads the 41 x 2 chain files PyHillTemp wrote, evaluates the temperature-1 log-likelihood of every row in ONE
batched GPU call per file (phf_log_target_batch) instead of a Python loop over rows
(compute_bayes_factors.py:11-27), integrates with the trapezium rule (:83) and writes BFs/<drug>_<channel>_B12.txt
(:86-100).  `pyhillfit_b200.ti.run_ti` is the fused alternative that never materialises the chains.
"""
import argparse
import sys

import numpy as np


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("-nc", "--num-cores", type=int, help="accepted for compatibility", default=1)
    # (not in the reference) the fused alternative for every pair of the data file at once: no chain files are read or
    # written -- pyhillfit_b200.ti.run_ti samples all (pair, model, temperature) chains on the device, accumulates the
    # temperature-1 log-likelihood in the kernel and integrates; writes the same BFs/<drug>_<channel>_B12.txt files
    parser.add_argument("--all-fused", action='store_true', default=False,
                        help="whole thermodynamic-integration sweep for every pair on the device (no chain files)")
    parser.add_argument("-i", "--iterations", type=int, default=500000, help="--all-fused: iterations per chain")
    parser.add_argument("-t", "--thinning", type=int, default=5, help="--all-fused: thinning")
    parser.add_argument("--seed", type=int, default=1, help="--all-fused: Philox seed")
    requiredNamed = parser.add_argument_group('required arguments')
    requiredNamed.add_argument("-d", "--drug", type=int, help="drug index (not needed with --all-fused)")
    requiredNamed.add_argument("-c", "--channel", type=int, help="channel index (not needed with --all-fused)")
    requiredNamed.add_argument("--data-file", type=str, required=True)
    return parser


def run_all_fused(dr, args):
    """Every (drug, channel) pair of the data file: ti.run_ti over the reference's ladder, then one B12 file per pair
    (python/compute_bayes_factors.py:83-100).  Under torchrun the sweep is sharded over the ranks; rank 0 writes."""
    import itertools as it
    from . import chainio
    from . import dist as phf_dist
    from .ti import run_ti
    jobs, data = [], []
    for top_drug, top_channel in it.product(dr.drugs, dr.channels):
        try:
            num_expts, experiment_numbers, experiments = dr.load_crumb_data(top_drug, top_channel)
        except Exception:
            continue
        concs = np.concatenate([experiments[i][:, 0] for i in range(num_expts)])
        responses = np.concatenate([experiments[i][:, 1] for i in range(num_expts)])
        if np.any(np.isnan(responses)):
            continue
        drug, channel, chain_file, images_dir = dr.nonhierarchical_chain_file_and_figs_dir(1, top_drug, top_channel, 1)
        jobs.append((drug, channel))
        data.append((concs, responses))
    phf_dist.init_process_group()          # (a no-op for a single process)
    res = run_ti(data, iterations=args.iterations, thinning=args.thinning, seed=args.seed)
    if phf_dist.world()[1] == 0:
        for (drug, channel), Bij in zip(jobs, res["B12"]):
            chainio.save_bayes_factor(drug, channel, Bij)
        print("{} Bayes factors written to BFs/".format(len(jobs)))
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        td.barrier()
        td.destroy_process_group()
    return 0


def compute_log_py_approxn(dr, model, pack, chain_file):
    from .sampler import log_target_batch
    chain = np.loadtxt(chain_file, usecols=range(dr.num_params), ndmin=2)
    _, l1 = log_target_batch(model, pack, chain, 0, 1.0)
    total = float(l1.sum().item())
    answer = total / chain.shape[0]
    if answer == -np.inf:
        print("ANSWER IS -INF")
    return answer


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) == 0:
        parser.print_help()
        return 1
    args = parser.parse_args(argv)
    from . import chainio
    from . import doseresponse as dr
    from .packing import SinglePack
    dr.setup(args.data_file)
    if args.all_fused:
        return run_all_fused(dr, args)
    if args.drug is None or args.channel is None:
        parser.error("the following arguments are required: -d/--drug, -c/--channel")
    top_drug, top_channel = dr.drugs[args.drug], dr.channels[args.channel]
    num_expts, experiment_numbers, experiments = dr.load_crumb_data(top_drug, top_channel)
    concs = np.concatenate([experiments[i][:, 0] for i in range(num_expts)])
    responses = np.concatenate([experiments[i][:, 1] for i in range(num_expts)])
    pack = SinglePack([(concs, responses)])
    expectations = {}
    for m in (1, 2):
        dr.define_model(m)
        temps = (np.arange(dr.n + 1.) / dr.n) ** dr.c
        log_p_ys = np.zeros(len(temps))
        for i, temp in enumerate(temps):
            print(temp)
            drug, channel, chain_file, images_dir = dr.nonhierarchical_chain_file_and_figs_dir(m, top_drug, top_channel, temp)
            log_p_ys[i] = compute_log_py_approxn(dr, m, pack, chain_file)
        print(log_p_ys)
        expectations[m] = dr.trapezium_rule(temps, log_p_ys)
        print(expectations)
    drug, channel, chain_file, images_dir = dr.nonhierarchical_chain_file_and_figs_dir(1, top_drug, top_channel, 1)
    Bij = np.exp(expectations[1] - expectations[2])
    chainio.save_bayes_factor(drug, channel, Bij)
    return 0


if __name__ == "__main__":
    sys.exit(main())
