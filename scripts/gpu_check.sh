#!/bin/bash
# developer tool (run under gpurun): GPU parity tests for the single-level path, the lane sweep and a short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -5
python scripts/sweep.py 2 64,1024 1:3,2:3 2>gpurun_out/err.txt | tail -4
python scripts/sweep.py 1 64 2:3,4:3 2>>gpurun_out/err.txt | tail -2
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs 2>>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench value %.4e  ms/step %.3f  roofline %.3f  kernel_ms %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('kernel_ms')))
"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --layout row --e2e-layout row 2>>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('row-major: bench value %.4e  ms/step %.3f  e2e %.4e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --e2e-layout chain 2>>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chain-major: bench value %.4e  ms/step %.3f  e2e %.4e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
"
