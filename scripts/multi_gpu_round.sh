#!/bin/bash
# developer tool (run under gpurun --gpus N): full-size config 5 and the thermodynamic-integration sweep on N GPUs
N=${1:-8}
TAG=${2:-r01_v8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR scripts/config5_full.py 1000000 2000 3 > $O/${TAG}_config5_${N}gpu.json 2> $O/${TAG}_config5_${N}gpu.err; tail -c 1500 $O/${TAG}_config5_${N}gpu.json
timeout 300 $TR scripts/full_ti_sweep.py 500000 8 > $O/${TAG}_ti_sweep_${N}gpu_8rep.txt 2> $O/${TAG}_ti_${N}gpu.err; cat $O/${TAG}_ti_sweep_${N}gpu_8rep.txt
timeout 300 $TR scripts/full_ti_sweep.py 500000 1 > $O/${TAG}_ti_sweep_${N}gpu_1rep.txt 2>> $O/${TAG}_ti_${N}gpu.err; cat $O/${TAG}_ti_sweep_${N}gpu_1rep.txt
tail -n 3 $O/${TAG}_config5_${N}gpu.err $O/${TAG}_ti_${N}gpu.err; exit 0
