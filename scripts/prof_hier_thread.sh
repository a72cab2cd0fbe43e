#!/bin/bash
# developer tool (run under gpurun): ncu --set full of the thread-per-chain hierarchical kernel at config 3's occupancy
TAG=${1:-r01_v7}
python scripts/prof_hier.py 3 256 500 > gpurun_out/${TAG}_hier_thread_plain.log 2>&1 || { tail -3 gpurun_out/${TAG}_hier_thread_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:am_hier_thread_kernel -s 1 -c 1 -o gpurun_out/${TAG}_hier_thread -f \
    python scripts/prof_hier.py 3 256 500 > gpurun_out/${TAG}_ncu_hier_thread.log 2>&1
tail -2 gpurun_out/${TAG}_hier_thread_plain.log; ls -la gpurun_out/${TAG}_hier_thread*
