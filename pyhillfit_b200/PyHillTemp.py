"""PyHillTemp command line (python/PyHillTemp.py of the reference): power-posterior chains on the 41-point
temperature ladder for one (drug index, channel index), all temperatures in one fused GPU launch.

Same flags as the reference (note: here -c is the CHANNEL index and -nc the core count, as in the reference).
Writes one header-less chain file per temperature with the float temperature in the path (PyHillTemp.py:165-169).
New optional flags: --num-chains (replicates per temperature), --seed, --segment.
"""
import argparse
import sys
import time

import numpy as np


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--iterations", type=int, help="number of MCMC iterations", default=500000)
    parser.add_argument("-t", "--thinning", type=int, help="how often to thin the MCMC, i.e. save every t-th iteration", default=5)
    parser.add_argument("-b", "--burn-in-fraction", type=int, help="given N saved MCMC iterations, discard the first N/b as burn-in", default=4)
    parser.add_argument("-a", "--all", action='store_true', default=False)
    parser.add_argument("-nc", "--num-cores", type=int, help="accepted for compatibility; all temperatures run in one GPU launch", default=1)
    parser.add_argument("-Ne", "--num_expts", type=int, help="how many experiments to fit to", default=0)
    parser.add_argument("--num-APs", type=int, default=500)
    parser.add_argument("--single", action='store_true', default=True)
    parser.add_argument("--hierarchical", action='store_true', default=False)
    parser.add_argument("--fix-hill", action='store_true', default=False)
    parser.add_argument("-bfo", "--best-fit-only", action='store_true', default=False)
    parser.add_argument("--num-chains", type=int, default=1, help="replicate chains per temperature [new]")
    parser.add_argument("--seed", type=int, default=1, help="Philox seed (the reference seeds numpy with 1) [new]")
    parser.add_argument("--segment", type=int, default=50000, help="iterations per kernel launch [new]")
    requiredNamed = parser.add_argument_group('required arguments')
    requiredNamed.add_argument("--data-file", type=str, required=True)
    requiredNamed.add_argument("-m", "--model", type=int, required=True)
    requiredNamed.add_argument("-d", "--drug", type=int, help="drug index", required=True)
    requiredNamed.add_argument("-c", "--channel", type=int, help="channel index", required=True)
    return parser


def do_mcmc_all_temperatures(dr, args, concs, responses, temperatures):
    """python/PyHillTemp.py:57-125 for every temperature (x replicates) at once -> [T, R, rows, d+1] host array
    of post-burn rows, plus the in-kernel mean temperature-1 log-likelihoods [T, R]."""
    import torch
    from .packing import SinglePack
    from .sampler import SingleLevelSampler
    R = args.num_chains
    T = len(temperatures)
    d = dr.num_params
    num_saved = args.iterations // args.thinning + 1
    burn = num_saved // args.burn_in_fraction
    pack = SinglePack([(concs, responses)])
    tt = np.repeat(np.asarray(temperatures, dtype=np.float64), R)
    # models 1 and 2 are run by separate invocations with the same --seed: the model goes into the Philox chain id
    # (as pyhillfit_b200/ti.py does) so that chain k of model 1 and chain k of model 2 do not share their draws
    s = SingleLevelSampler(args.model, pack, np.zeros(T * R, dtype=np.int32), tt, np.ones((T * R, d)), variant="temp",
                           seed=args.seed, chain_id_base=(args.model - 1) * (1 << 40), thinning=args.thinning,
                           burn_rows=burn)
    # only the rows the reference keeps (chain[burn:], PyHillTemp.py:125) are written by the kernel and copied back
    kept = num_saved - max(burn, 0)
    chain = torch.empty((s.n, kept, d + 1), dtype=torch.float64, device=s.device)
    at = 0
    if burn == 0:
        chain[:, 0, :] = s.initial_row()
        at = 1
    done = 0
    while done < args.iterations:
        k = min(args.segment - args.segment % args.thinning or args.thinning, args.iterations - done)
        seg = s.run(k, discard_burn=True)
        chain[:, at:at + seg.shape[1], :] = seg
        at += seg.shape[1]
        done += k
    torch.cuda.synchronize()
    assert at == kept
    host = chain.cpu().numpy().reshape(T, R, kept, d + 1)
    return host, s.loglik_t1_mean().reshape(T, R)


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) == 0:
        parser.print_help()
        return 1
    args = parser.parse_args(argv)
    from . import chainio
    from . import doseresponse as dr
    dr.define_model(args.model)
    dr.setup(args.data_file)
    drug, channel = dr.drugs[args.drug], dr.channels[args.channel]
    num_expts, experiment_numbers, experiments = dr.load_crumb_data(drug, channel)
    concs = np.concatenate([experiments[i][:, 0] for i in range(num_expts)])
    responses = np.concatenate([experiments[i][:, 1] for i in range(num_expts)])
    temperatures = (np.arange(dr.n + 1.) / dr.n) ** dr.c
    print("\nDoing temperatures: {}\n".format(temperatures))
    start = time.time()
    chains, ll1 = do_mcmc_all_temperatures(dr, args, concs, responses, temperatures)
    print("\nMCMC time: {} s\n".format(int(time.time() - start)))
    for i, temperature in enumerate(temperatures):
        cdrug, cchannel, chain_file, images_dir = dr.nonhierarchical_chain_file_and_figs_dir(args.model, drug, channel, temperature)
        print("chain_file:", chain_file)
        for r in range(args.num_chains):
            f = chain_file if r == 0 else chainio.extra_chain_name(chain_file, r)
            chainio.save_tempered_chain(f, chains[i, r])
    return 0


if __name__ == "__main__":
    sys.exit(main())
