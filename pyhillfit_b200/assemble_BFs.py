"""assemble_BFs command line (python/assemble_BFs.py of the reference): the bookkeeping pass over the files the path
writes.  Reads BFs/<drug>_<channel>_B12.txt (python/assemble_BFs.py:55-58) and both models' best_fit_params.txt
(:60-63) for every pair of the data file and prints how many Bayes factors fall into each evidence band (:94-120:
"substantial" 3-10, "strong" 10-100, "decisive" > 100, either way round, and "no evidence" between 1/3 and 3).  Host
only; nothing here is on the hot path -- it exists so that the files written by PyHillFit --best-fit-only and
compute_bayes_factors (per pair or --all-fused) are exercised by the consumer the reference has for them.  The
reference's interactive figure browsing after its sys.exit() (:122 on) is unreachable code there and is not mirrored.
"""
import argparse
import itertools as it
import os
import sys

import numpy as np

BANDS = ("substantial_b12", "strong_b12", "decisive_b12", "substantial_b21", "strong_b21", "decisive_b21")


def evidence_band(b12):
    """The band of one Bayes factor, or None (python/assemble_BFs.py:94-105; boundaries as there: (3, 10], (10, 100],
    (100, inf) on B12, then the same on 1/B12)."""
    for value, suffix in ((b12, "b12"), (1.0 / b12 if b12 != 0 else np.inf, "b21")):
        if 3 < value <= 10:
            return "substantial_" + suffix
        if 10 < value <= 100:
            return "strong_" + suffix
        if 100 < value:
            return "decisive_" + suffix
    return None


def summarise(b12s):
    """Counts per band + "no_evidence" (1/3 < B12 < 3, :107-108) + "ambiguous" (0.9 < B12 < 1.1, :73) for an
    iterable of Bayes factors."""
    counts = {k: 0 for k in BANDS}
    counts["no_evidence"] = counts["ambiguous"] = 0
    for b in b12s:
        band = evidence_band(b)
        if band:
            counts[band] += 1
        if 1. / 3 < b < 3.:
            counts["no_evidence"] += 1
        if 0.9 < b < 1.1:
            counts["ambiguous"] += 1
    return counts


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--skip-missing", action='store_true', default=False,
                        help="(not in the reference) skip pairs whose files are absent instead of failing")
    requiredNamed = parser.add_argument_group('required arguments')
    requiredNamed.add_argument("--data-file", type=str, required=True,
                               help="csv file from which to read in data, in same format as provided crumb_data.csv")
    return parser


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) == 0:
        parser.print_help()
        return 1
    args = parser.parse_args(argv)
    from . import doseresponse as dr
    dr.setup(args.data_file)
    BFs, best_params = {}, ({}, {})
    for top_drug, top_channel in it.product(dr.drugs, dr.channels):
        drug, channel, chain_file, images_dir = dr.nonhierarchical_chain_file_and_figs_dir(1, top_drug, top_channel, 1)
        bf_file = "BFs/{}_{}_B12.txt".format(drug, channel)
        if args.skip_missing and not os.path.exists(bf_file):
            continue
        BFs[(top_drug, top_channel)] = float(np.loadtxt(bf_file))
        for m in (1, 2):
            drug, channel, chain_file, images_dir = dr.nonhierarchical_chain_file_and_figs_dir(m, top_drug, top_channel, 1)
            f = images_dir + "{}_{}_best_fit_params.txt".format(drug, channel)
            if os.path.exists(f) or not args.skip_missing:
                best_params[m - 1][(top_drug, top_channel)] = np.loadtxt(f)
    counts = summarise(BFs.values())
    print("NO EVIDENCE:", counts["no_evidence"])
    for k in BANDS:
        print(k + ":", counts[k])
    return 0


if __name__ == "__main__":
    sys.exit(main())
