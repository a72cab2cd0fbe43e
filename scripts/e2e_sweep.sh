#!/bin/bash
# developer tool: e2e throughput vs number of overlapped segments
python scripts/d2h_probe.py
for sgm in 8 16 32 64; do
  python bench.py --steps 5 --no-other-configs --no-cpu-baseline --e2e-segments $sgm 2>/dev/null > /tmp/b.json
  python -c "import json; d=json.load(open('/tmp/b.json')); print('segments $sgm value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
done
