#!/bin/bash
# developer tool (run under gpurun, ONE GPU): what the driver runs at round end, plus the bench evidence
# usage: round_check.sh <tag>     (writes gpurun_out/<tag>_*)
TAG=${1:-r01_v8}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} ) > $O/${TAG}_pytest_gpu.log 2>&1
tail -4 $O/${TAG}_pytest_gpu.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $O/${TAG}_smoke.log 2>&1
tail -4 $O/${TAG}_smoke.log
( time timeout 500 python bench.py > $O/${TAG}_bench.json ) 2> $O/${TAG}_bench.err || tail -5 $O/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("bench value %.4e e2e %.4e roofline %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
    print("config3", json.dumps(d["other_configs"]["config3_hierarchical"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json ) 2> $O/${TAG}_bench_reference.err
tail -c 600 $O/${TAG}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > $O/${TAG}_ncu_bench.log 2>&1
ls -la $O/${TAG}_*
