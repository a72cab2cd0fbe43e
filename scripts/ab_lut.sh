#!/bin/bash
# developer tool (run under gpurun, ONE GPU): A/B of two library builds (PHF_B200_LIB) on the probes and the bench
# usage: ab_lut.sh <tag> <libA> <libB>
TAG=${1:-ab}
A=${2:-pyhillfit_b200/libphf_b200.so}
B=${3:-pyhillfit_b200/libphf_b200_nolut.so}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
for L in $A $B; do
  N=$(basename $L .so)
  export PHF_B200_LIB=$PWD/$L
  echo "== $N" | tee -a $O/${TAG}_probes.txt
  timeout 200 python scripts/occupancy_probe.py 2 2 2>&1 | tee -a $O/${TAG}_probes.txt
  timeout 200 python scripts/occupancy_probe.py 2 1 2>&1 | tail -4 | tee -a $O/${TAG}_probes.txt
  timeout 200 python scripts/hier_probe.py 3 256 2>&1 | tail -3 | tee -a $O/${TAG}_probes.txt
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_$N.json 2> $O/${TAG}_bench_$N.err
  python - <<PY | tee -a $O/${TAG}_probes.txt
import json
d = json.loads(open("$O/${TAG}_bench_$N.json").read().strip().splitlines()[-1])
o = d["other_configs"]
print("bench value %.4e e2e %.4e roofline %.3f kernel_ms %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["kernel_ms"]))
print("config3 %.4e  config4 %.4e  config5 %.4e" % (o["config3_hierarchical"]["value"], o["config4_ti_64_temperatures"]["value"], o["config5_synthetic_share"]["value"]))
PY
done
