#!/bin/bash
PTB_BATCH=8388608 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
PTB_BATCH=8388608 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 0 -c 4 \
   -f -o gpurun_out/trace_v7 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v7.log 2>&1
echo rc=$?
