#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for r in 8 14; do
  PTB_REFILL=$r python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c_refill_$r.json 2>gpurun_out/c_refill_$r.err
done
python scripts/dev_check.py > gpurun_out/dev_check4.log 2>&1
echo done
