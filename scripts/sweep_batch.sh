#!/bin/bash
# PTB_BATCH sweep on the headline scene (spp 64): step, trace and the rest (k_shade mostly) per frame
for b in ${BATCHES:-8388608 33554432 67108864 134217728 268435456}; do
  PTB_BATCH=$b python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); t=d['roofline']['trace_ms_per_step']; print('batch $b', 'Mpaths/s', round(d['value']), 'step ms', round(d['ms_per_step'],2), 'trace ms', round(t,2), 'rest ms', round(d['ms_per_step']-t,2), 'launches', d['gpu_launches'])"
done
