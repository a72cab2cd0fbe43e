#!/bin/bash
# developer tool (run under gpurun --gpus 8): bench.py at N = 8 and 4 exactly as the driver launches it (final kernels)
O=gpurun_out
TAG=${1:-r02_v3}
for n in 8 4 2; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > $O/${TAG}_bench_${n}gpu.json 2> $O/${TAG}_bench_${n}gpu.err || tail -5 $O/${TAG}_bench_${n}gpu.err
done
python - <<PY
import json
for n in (8, 4, 2):
    try:
        b = json.loads(open("$O/${TAG}_bench_%dgpu.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4e e2e %.4e" % (b["value"], b["e2e"]["value"]), {k: ("%.3e" % v["value"], round(v["seconds"], 4), v.get("speculation_rank0"), v.get("rank_count_independence", {}).get("B12_sha256_16")) for k, v in b["other_configs"].items()})
    except Exception as e:
        print(n, "unreadable", e)
PY
