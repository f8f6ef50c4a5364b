#!/bin/bash
# baseline of the current code: gpu tests, refill sweep, ncu capture (source-level) of k_trace/k_shade
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for r in 8 14 20 26; do
  PTB_REFILL=$r python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_refill_$r.json 2>/dev/null
done
PTB_BATCH=8388608 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
PTB_BATCH=8388608 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_raygen' -s 0 -c 5 \
   -f -o gpurun_out/trace_v4base python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v4base.log 2>&1
echo done
