for T in 1 2; do for seg in 32 64; do
  timeout 150 python bench.py --steps 3 --warmup 3 --no-strong --no-other-configs --no-cpu-baseline --e2e-segments $seg --e2e-steps 12 --e2e-threads-per-model $T 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads/model $T segments $seg e2e %.4e value %.4e' % (b['e2e']['value'], b['value']))"
done; done
