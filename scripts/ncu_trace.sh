#!/bin/bash
# one `ncu --set full` capture of the first k_trace / k_shade launches of an 8 Mi-path batch (bounce 0 and 1)
# usage: gpurun -- bash scripts/ncu_trace.sh <tag>
tag=${1:-cap}
mkdir -p gpurun_out
PTB_BATCH=8388608 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_plain.log 2>&1 &&
PTB_BATCH=8388608 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 0 -c 4 \
   -f -o gpurun_out/${tag} python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep
