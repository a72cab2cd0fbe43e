#!/bin/bash
# developer tool (run under gpurun, ONE GPU): the evidence set committed under profiles/ -- bench line, ncu launch
# list of the bench command, ncu --set full captures of the sampler kernels at the occupancies the bench runs them at.
# usage: profile_round.sh <tag>     (writes gpurun_out/<tag>_*)
TAG=${1:-r01_v7}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { tail -5 $O/${TAG}_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > $O/${TAG}_ncu_bench.log 2>&1
# model 2 at the config-2 occupancy (26 880 chains of model 2 alone, 2 lanes per chain) and at one thread per chain
ncu --set full --clock-control none --import-source on -k regex:am_single_kernel -s 2 -c 1 -o $O/${TAG}_g2 -f \
    python scripts/prof_run.py 128 500 2 2 > $O/${TAG}_ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:am_single_kernel -s 2 -c 1 -o $O/${TAG}_g1 -f \
    python scripts/prof_run.py 1024 500 2 1 > $O/${TAG}_ncu_g1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:am_hier_kernel -s 2 -c 1 -o $O/${TAG}_hier3 -f \
    python scripts/prof_hier.py 3 256 500 16 > $O/${TAG}_ncu_hier.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:am_hier_thread_kernel -s 2 -c 1 -o $O/${TAG}_hier_thread -f \
    python scripts/prof_hier.py 3 256 500 1 > $O/${TAG}_ncu_hier_thread.log 2>&1
ls -la $O/${TAG}_*
