#!/bin/bash
# developer tool (run under gpurun, ONE GPU): chain blocks by decreasing cost vs index order
O=gpurun_out; TAG=${1:-ab2}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
for ORD in 0 1 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --cta-order $ORD > $O/${TAG}_bench_order$ORD.json 2> $O/${TAG}_bench_order$ORD.err
  python - <<PY | tee -a $O/${TAG}_order.txt
import json
d = json.loads(open("$O/${TAG}_bench_order$ORD.json").read().strip().splitlines()[-1])
print("cta_order $ORD: bench value %.4e ms/step %.3f roofline %.3f kernel_ms %s on_stream %s launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms"], d["roofline"]["launch_ms_on_stream"], d["gpu_launches"]))
PY
done
