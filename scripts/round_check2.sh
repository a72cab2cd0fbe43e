#!/bin/bash
# developer tool (run under gpurun, ONE GPU): what the driver runs at round end -- GPU tests, smoke(), both bench arms.
# usage: round_check2.sh <tag>     (writes gpurun_out/<tag>_*)
TAG=${1:-r02_v2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${TAG}_smoke.log 2>&1
tail -4 $O/${TAG}_smoke.log
( time timeout 600 python bench.py > $O/${TAG}_bench.json ) 2> $O/${TAG}_bench.err || tail -5 $O/${TAG}_bench.err
( time timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json ) 2> $O/${TAG}_bench_reference.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench value %.4e e2e %.4e roofline %.3f ess/s %.3e launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ess_per_s"], d["gpu_launches"]))
print({k: ("%.3e" % v["value"], round(v["roofline"]["frac"], 3)) for k, v in d["other_configs"].items()})
r = json.loads(open("$O/${TAG}_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm %.4e ess/s %.3e" % (r["value"], r["ess_per_s"]), "config equal:", r["config"] == {k: v for k, v in d["config"].items() if k != "l2"})
PY
grep real $O/${TAG}_bench.err $O/${TAG}_bench_reference.err
