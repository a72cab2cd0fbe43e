#!/bin/bash
# developer demo (GPU box, repo root): the all-pairs Bayes-factor story at the reference's default sizes, timed:
# least-squares fits of both models, the fused thermodynamic-integration sweep (210 pairs x 2 models x 41 temperatures x
# 500 000 iterations), the evidence bands.  usage: cli_demo_fused.sh [n_gpus]
N=${1:-1}
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
for m in 1 2; do
  echo "== PyHillFit -a -m $m --best-fit-only"; S=$(date +%s%N)
  python -m pyhillfit_b200.PyHillFit --data-file data/crumb_data.csv -m $m -a --best-fit-only > run.log 2>&1 || tail -3 run.log
  echo "wall $(( ($(date +%s%N) - S) / 1000000 )) ms"
done
echo "== compute_bayes_factors --all-fused on $N GPU(s)"; S=$(date +%s%N)
if [ "$N" = "1" ]; then
  python -m pyhillfit_b200.compute_bayes_factors --data-file data/crumb_data.csv --all-fused > run.log 2>&1 || tail -3 run.log
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      -m pyhillfit_b200.compute_bayes_factors --data-file data/crumb_data.csv --all-fused > run.log 2>&1 || tail -3 run.log
fi
tail -1 run.log
echo "wall $(( ($(date +%s%N) - S) / 1000000 )) ms"
ls BFs | wc -l; cat BFs/Amiodarone_hERG_B12.txt
echo "== assemble_BFs"; python -m pyhillfit_b200.assemble_BFs --data-file data/crumb_data.csv
cd /; rm -rf $W
