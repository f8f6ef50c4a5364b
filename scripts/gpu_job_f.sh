#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for cfg in "1 14" "6 14" "8 14" "12 14" "8 8" "8 20"; do
  set -- $cfg
  PTB_LEAFMIN=$1 PTB_REFILL=$2 python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/f_$1_$2.json 2>gpurun_out/f.err
  python - <<PY
import json
try:
  d=json.load(open('gpurun_out/f_$1_$2.json'))
  print('leafmin $1 refill $2', 'Mpaths/s %.0f ms/step %.1f trace_ms %.1f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['trace_ms_per_step'],d['roofline']['frac']))
except Exception as e: print('$1 $2','ERR',e)
PY
done
