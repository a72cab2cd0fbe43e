#!/bin/bash
# developer demo: hierarchical fit of the reference's synthetic_data.csv (groups of 5, 5 and 50 experiments)
ROOT=$(pwd); W=$(mktemp -d); cd $W; mkdir data
PYTHONPATH=$ROOT:$ROOT/tests python - <<'PY'
import numpy as np, os
from _data import GOLD
z = np.load(os.path.join(GOLD, "datasets.npz"))
with open("data/synthetic_data.csv", "w") as out:
    out.write("Compound,Channel,Experiment,Dose,Response\n")
    for row in zip(z["synthetic_data__drug"], z["synthetic_data__channel"], z["synthetic_data__experiment"], z["synthetic_data__dose"], z["synthetic_data__response"]):
        out.write("%s,%s,%d,%r,%r\n" % (row[0], row[1], row[2], float(row[3]), float(row[4])))
PY
export PYTHONPATH=$ROOT
S=$SECONDS
python -m pyhillfit_b200.PyHillFit --data-file data/synthetic_data.csv -m 2 -a --hierarchical -i ${1:-20000} --num-chains 4 > run.log 2>&1
grep -E "hierarchical chains|Traceback|Error" run.log; tail -3 run.log
echo "wall $((SECONDS-S)) s"; find output -name "*chain.txt" | head; python - <<'PY'
import numpy as np, glob
for f in sorted(glob.glob("output/*/hierarchical/*/*/*_expts/chain/*chain.txt")):
    c = np.loadtxt(f); b = len(c)//4
    print(f.split("/")[-1], c.shape, "alpha %.3f beta %.2f mu %.3f s %.3f sigma %.2f" % tuple(np.median(c[b:, [0,1,2,3,-2]], axis=0)))
PY
cd /; rm -rf $W
