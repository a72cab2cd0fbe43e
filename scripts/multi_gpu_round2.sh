#!/bin/bash
# developer tool (run under gpurun --gpus 8): the strong / weak scaling evidence of round 2 -- bench.py at N = 8, 4, 2
# exactly as the driver launches it, and the D2H-only host ceiling at N = 1, 2, 4, 8.
O=gpurun_out
TAG=${1:-r02_v1}
for n in 8 4 2; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > $O/${TAG}_bench_${n}gpu.json 2> $O/${TAG}_bench_${n}gpu.err || tail -5 $O/${TAG}_bench_${n}gpu.err
done
: > $O/${TAG}_d2h_probe_ranks.txt
for n in 1 2 4 8; do for pin in 0 1; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      scripts/d2h_probe_ranks.py $pin 2>/dev/null | grep "D2H only" >> $O/${TAG}_d2h_probe_ranks.txt
done; done
nproc >> $O/${TAG}_d2h_probe_ranks.txt; numactl -H 2>/dev/null | head -5 >> $O/${TAG}_d2h_probe_ranks.txt; lscpu | grep -E "NUMA|Model name|Socket" >> $O/${TAG}_d2h_probe_ranks.txt
cat $O/${TAG}_d2h_probe_ranks.txt
python - <<PY
import json
for n in (8, 4, 2):
    try:
        b = json.loads(open("$O/${TAG}_bench_%dgpu.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4e e2e %.4e" % (b["value"], b["e2e"]["value"]), {k: ("%.3e" % v["value"], round(v["seconds"], 4), v.get("lanes_rank0"), v.get("rank_count_independence", {}).get("B12_sha256_16")) for k, v in b["other_configs"].items()})
    except Exception as e:
        print(n, "unreadable", e)
PY
