#!/bin/bash
# PTB_REFILL sweep on the headline scene (spp 64): trace ms per frame
for r in 8 10 12 14 16 18; do
  PTB_REFILL=$r python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('refill $r', 'Mpaths/s', round(d['value']), 'trace ms', round(d['roofline']['trace_ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))"
done
