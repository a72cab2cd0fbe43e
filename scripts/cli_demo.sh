#!/bin/bash
# developer demo: the reference's command lines at their default sizes, timed (run from the repo root on a GPU box)
ROOT=$(pwd)
W=$(mktemp -d); cd $W; mkdir data
PYTHONPATH=$ROOT:$ROOT/tests python - <<'PY'
import numpy as np, os, sys
from _data import GOLD
z = np.load(os.path.join(GOLD, "datasets.npz"))
with open("data/crumb_data.csv", "w") as out:
    out.write("Compound,Channel,Experiment,Dose,Response\n")
    for row in zip(z["crumb_data__drug"], z["crumb_data__channel"], z["crumb_data__experiment"], z["crumb_data__dose"], z["crumb_data__response"]):
        out.write("%s,%s,%d,%r,%r\n" % (row[0], row[1], row[2], float(row[3]), float(row[4])))
PY
export PYTHONPATH=$ROOT
echo "== PyHillFit -a -m 2 (210 pairs, 500000 iterations each)"; S=$SECONDS; python -m pyhillfit_b200.PyHillFit --data-file data/crumb_data.csv -m 2 -a > run.log 2>&1; grep -E "chains x|wall| s$" run.log; tail -2 run.log | grep -i -E "error|Traceback"; echo "wall $((SECONDS-S)) s"
du -sh output | tail -1; ls output/crumb_data/single-level | wc -l
echo "== PyHillFit -a --hierarchical (210 pairs)"; S=$SECONDS; python -m pyhillfit_b200.PyHillFit --data-file data/crumb_data.csv -m 2 -a --hierarchical > run.log 2>&1; grep -E "hierarchical chains|wall" run.log; tail -2 run.log | grep -i -E "error|Traceback"; echo "wall $((SECONDS-S)) s"
echo "== PyHillTemp (41 temperatures) x 2 models + compute_bayes_factors, drug 0 channel 0"
S=$SECONDS; python -m pyhillfit_b200.PyHillTemp --data-file data/crumb_data.csv -m 1 -d 0 -c 0 > run.log 2>&1; grep -E "MCMC time|wall" run.log; tail -2 run.log | grep -i -E "error|Traceback"; echo "wall $((SECONDS-S)) s"
S=$SECONDS; python -m pyhillfit_b200.PyHillTemp --data-file data/crumb_data.csv -m 2 -d 0 -c 0 > run.log 2>&1; grep -E "MCMC time|wall" run.log; tail -2 run.log | grep -i -E "error|Traceback"; echo "wall $((SECONDS-S)) s"
S=$SECONDS; python -m pyhillfit_b200.compute_bayes_factors --data-file data/crumb_data.csv -d 0 -c 0 > run.log 2>&1; grep -E "wall" run.log; tail -2 run.log | grep -i -E "error|Traceback"; echo "wall $((SECONDS-S)) s"
cat BFs/Amiodarone_hERG_B12.txt
echo "== construct_hierarchical_cdfs, drug 1 channel 1"; S=$SECONDS; python -m pyhillfit_b200.construct_hierarchical_cdfs --data-file data/crumb_data.csv --selection 1:1 > run.log 2>&1; grep -E "done|wall" run.log; tail -2 run.log | grep -i -E "error|Traceback"; echo "wall $((SECONDS-S)) s"
du -sh output | tail -1
cd /; rm -rf $W
