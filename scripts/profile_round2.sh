#!/bin/bash
# developer tool (run under gpurun, ONE GPU): the round-2 evidence set committed under profiles/ -- bench line, reference
# arm, ncu launch list of the bench command, ncu --set full captures of the kernels at the occupancies they run at.
# Every ncu pass profiles a command that has already exited 0 without ncu.  usage: profile_round2.sh <tag>
TAG=${1:-r02_v1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { tail -5 $O/${TAG}_bench.err; exit 1; }
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > $O/${TAG}_ncu_bench.log 2>&1
cap() {  # name, kernel regex, command...
    local name=$1 rx=$2; shift 2
    timeout 200 "$@" > $O/${TAG}_run_${name}.log 2>&1 || { echo "$name: plain run failed"; tail -3 $O/${TAG}_run_${name}.log; return; }
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o $O/${TAG}_${name} -f "$@" > $O/${TAG}_ncu_${name}.log 2>&1
    ncu -i $O/${TAG}_${name}.ncu-rep --page raw --csv > $O/${TAG}_${name}_raw.csv 2>/dev/null
}
# model 2 at the config-2 occupancy (26 880 chains of model 2 alone, 2 lanes per chain); one thread per chain at 215 040
cap g2 am_single_kernel python scripts/prof_run.py 128 500 2 2
cap g1 am_single_kernel python scripts/prof_run.py 1024 500 2 1
# the speculative form at the sharded-sweep occupancy (2 100 chains, 2 lanes x 4 hypotheses)
cap spec24 am_single_spec_kernel python scripts/prof_run.py 10 2000 2 2 0 4
# hierarchical thread-per-chain kernel: Ne = 3 (39 424 chains) and Ne = 4 (10 496 chains); lane kernel Ne = 5
cap hier_thread3 am_hier_thread_kernel python scripts/prof_hier.py 3 256 500 1
cap hier_thread4 am_hier_thread_kernel python scripts/prof_hier.py 4 256 500 1
cap hier_lane5 am_hier_kernel python scripts/prof_hier.py 5 256 500 16
# the FP64 peak probe itself: the roofline's denominator (sm__inst_executed_pipe_fp64 of fp64_peak_kernel)
timeout 300 ncu --set full --clock-control none -k regex:fp64_peak_kernel -s 1 -c 1 -o $O/${TAG}_fp64_peak -f \
    python -c "from pyhillfit_b200 import _lib; print(_lib.fp64_peak_tflops(3))" > $O/${TAG}_ncu_fp64_peak.log 2>&1
ncu -i $O/${TAG}_fp64_peak.ncu-rep --page raw --csv > $O/${TAG}_fp64_peak_raw.csv 2>/dev/null
# dynamic opcode mixes (warp-instructions per warp-iteration)
python scripts/ncu_opmix.py $O/${TAG}_g2.ncu-rep $((53760 / 32 * 500)) > $O/${TAG}_g2_opmix.txt 2>&1
python scripts/ncu_opmix.py $O/${TAG}_g1.ncu-rep $((215040 / 32 * 500)) > $O/${TAG}_g1_opmix.txt 2>&1
python scripts/ncu_opmix.py $O/${TAG}_spec24.ncu-rep $((2100 * 8 / 32 * 2000)) > $O/${TAG}_spec24_opmix.txt 2>&1
python scripts/ncu_opmix.py $O/${TAG}_hier_thread3.ncu-rep $((39424 / 32 * 500)) > $O/${TAG}_hier_thread3_opmix.txt 2>&1
python scripts/ncu_opmix.py $O/${TAG}_hier_thread4.ncu-rep $((10496 / 32 * 500)) > $O/${TAG}_hier_thread4_opmix.txt 2>&1
timeout 120 python scripts/d2h_probe_ranks.py > $O/${TAG}_d2h_probe_1gpu.txt 2>&1
ls -la $O/${TAG}_* | awk '{print $5, $9}'
