#!/bin/bash
# developer tool (run under gpurun, ONE GPU): A/B of library builds (PHF_B200_LIB) on the probes and a short bench
# usage: ab_libs.sh <tag> <lib> [<lib> ...]      (the GPU tests run once, on the default library)
TAG=$1; shift
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
for L in "$@"; do
  N=$(basename $L .so)
  export PHF_B200_LIB=$PWD/$L
  echo "== $N" | tee -a $O/${TAG}_probes.txt
  timeout 200 python scripts/occupancy_probe.py 2 2 2>&1 | tail -3 | tee -a $O/${TAG}_probes.txt
  timeout 200 python scripts/hier_probe.py 3 256 2>&1 | tail -2 | tee -a $O/${TAG}_probes.txt
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_$N.json 2> $O/${TAG}_bench_$N.err
  python - <<PY | tee -a $O/${TAG}_probes.txt
import json
d = json.loads(open("$O/${TAG}_bench_$N.json").read().strip().splitlines()[-1])
o = d["other_configs"]
print("bench value %.4e roofline %.3f kernel_ms %s on_stream %s" % (d["value"], d["roofline"]["frac"], d["kernel_ms"], d["roofline"]["launch_ms_on_stream"]))
print("config3 %.4e  config4 %.4e  config5 %.4e" % (o["config3_hierarchical"]["value"], o["config4_ti_64_temperatures"]["value"], o["config5_synthetic_share"]["value"]))
PY
done
