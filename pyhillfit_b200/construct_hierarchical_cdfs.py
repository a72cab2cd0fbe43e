"""construct_hierarchical_cdfs command line (python/construct_hierarchical_cdfs.py of the reference): posterior-
predictive CDFs of Hill and pIC50 from a hierarchical chain file, and inverse-CDF samples for the action-potential
runs.  The per-row scipy.stats loop (:46-50) is one GPU reduction (phf_hier_predictive_cdfs).  Same flags; plots
are not drawn.  Files: <...>/cdfs/*_posterior_predictive_{hill,pic50}_cdf.txt and
posterior_predictive_hill_pic50_samples/*_hill_pic50_samples.txt, formats as the reference writes them (:128-149).
"""
import argparse
import itertools as it
import sys

import numpy as np

from . import _lib

NUM_X_PTS, HILL_RANGE, PIC50_RANGE = 501, (0., 4.), (-2., 12.)      # construct_hierarchical_cdfs.py:33-37


def construct_posterior_predictive_cdfs(alphas, betas, mus, ss):
    """Same return tuple as the reference function (:32-58):
    hill_x_range, hill_cdf, pic50_x_range, pic50_cdf, hill_pdf, pic50_pdf."""
    rows = np.ascontiguousarray(np.stack([alphas, betas, mus, ss], axis=1), dtype=np.float64)
    out = predictive_cdfs_from_rows(rows)
    hill_x_range = np.linspace(HILL_RANGE[0], HILL_RANGE[1], NUM_X_PTS)
    pic50_x_range = np.linspace(PIC50_RANGE[0], PIC50_RANGE[1], NUM_X_PTS)
    return hill_x_range, out[0], pic50_x_range, out[2], out[1], out[3]


def predictive_cdfs_from_rows(rows, device=None):
    """rows: host array or device tensor [n, >=4] whose first four columns are (alpha, beta, mu, s) -> host
    array [4, 501] = (hill cdf, hill pdf, pic50 cdf, pic50 pdf)."""
    torch = _lib.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = rows if isinstance(rows, torch.Tensor) else torch.from_numpy(np.array(rows, dtype=np.float64))
    t = t.to(device)
    if t.dim() != 2 or t.shape[1] < 4 or t.stride(1) != 1:
        t = t.reshape(-1, t.shape[-1]).contiguous()
    out = torch.empty((4, NUM_X_PTS), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().phf_hier_predictive_cdfs(t.shape[0], t.data_ptr(), t.stride(0), NUM_X_PTS,
                                                        HILL_RANGE[0], HILL_RANGE[1], PIC50_RANGE[0], PIC50_RANGE[1],
                                                        out.data_ptr(), _lib.current_stream_ptr()),
                   "phf_hier_predictive_cdfs")
    return out.cpu().numpy()


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("-s", "--samples", type=int, help="number of Hill and pIC50 samples for use in AP model", default=500)
    parser.add_argument("-a", "--all", action='store_true', default=False)
    parser.add_argument("--num-cores", type=int, help="accepted for compatibility", default=1)
    parser.add_argument("-np", "--no-plots", action='store_true', help="accepted for compatibility (no plots are drawn)", default=False)
    parser.add_argument("-tu", "--top-up", action='store_true', default=False)
    parser.add_argument("-sy", "--synthetic", action='store_true', default=False)
    parser.add_argument("-Ne", "--num_expts", type=int, help="how many experiments to fit to", default=0)
    parser.add_argument("--data-file", type=str, help="csv file from which to read in data, in same format as provided crumb_data.csv")
    parser.add_argument("--selection", type=str, default=None, help="'d1,d2:c1,c2' 1-based drug and channel numbers [new]")
    return parser


def run(dr, args, drug, channel, rng):
    print("\n\n{} + {}\n\n".format(drug, channel))
    num_expts, experiment_numbers, experiments = dr.load_crumb_data(drug, channel)
    if 0 < args.num_expts < num_expts:
        num_expts = args.num_expts
        save_samples_for_APs = False
    else:
        print("Fitting to all experiments\n")
        save_samples_for_APs = True
    drug, channel, output_dir, chain_dir, figs_dir, chain_file = dr.hierarchical_output_dirs_and_chain_file(drug, channel, num_expts)
    try:
        mcmc = np.loadtxt(chain_file, usecols=range(4))
    except IOError:
        print("tried loading", chain_file)
        print("No MCMC file found for {} + {}\n".format(drug, channel))
        return None
    burn = mcmc.shape[0] // 4
    mcmc = mcmc[burn:, :]
    hill_x, hill_cdf, pic50_x, pic50_cdf, hill_pdf, pic50_pdf = construct_posterior_predictive_cdfs(
        mcmc[:, 0], mcmc[:, 1], mcmc[:, 2], mcmc[:, 3])
    hill_cdf_file, pic50_cdf_file = dr.hierarchical_posterior_predictive_cdf_files(drug, channel, num_expts)
    np.savetxt(hill_cdf_file, np.vstack((hill_x, hill_cdf)).T)
    np.savetxt(pic50_cdf_file, np.vstack((pic50_x, pic50_cdf)).T)
    hill_samples = np.interp(rng.rand(args.samples), hill_cdf, hill_x)        # inverse-cdf sampling (:136-141)
    pic50_samples = np.interp(rng.rand(args.samples), pic50_cdf, pic50_x)
    if save_samples_for_APs:
        samples_file = dr.hierarchical_hill_and_pic50_samples_for_AP_file(drug, channel)
        with open(samples_file, 'w') as outfile:
            outfile.write('# {} samples of (Hill,pIC50) drawn from their posterior predictive distributions, as defined by MCMC samples\n'.format(args.samples))
            np.savetxt(outfile, np.vstack((hill_samples, pic50_samples)).T)
    print("\n{} + {} done!\n".format(drug, channel))
    return hill_cdf_file, pic50_cdf_file


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    from . import doseresponse as dr
    dr.setup(args.data_file)
    if args.selection:
        ds, cs = args.selection.split(":")
        drugs = [dr.drugs[int(x) - 1] for x in ds.split(",")]
        channels = [dr.channels[int(x) - 1] for x in cs.split(",")]
    else:
        drugs, channels = dr.list_drug_channel_options(args.all)
    rng = np.random.RandomState(1)                                            # npr.seed(1) at :13-14
    for drug, channel in it.product(drugs, channels):
        try:
            run(dr, args, drug, channel, rng)
        except Exception as e:
            print(e)
            print("Failed to run {} + {}!".format(drug, channel))
    return 0


if __name__ == "__main__":
    sys.exit(main())
