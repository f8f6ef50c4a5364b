#!/bin/bash
# k_shade time against blocks per SM, bulk copies on/off; usage: gpurun -- bash scripts/sweep_shade.sh
mkdir -p gpurun_out
run() { python bench.py --spp 128 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'Mpaths/s %.0f' % d['value'], 'shade ms %.2f' % d['roofline_shade']['ms_per_step'], 'frac %.3f' % d['roofline_shade']['frac'], 'trace ms %.2f' % d['roofline']['trace_ms_per_step'])"; }
PTB_SHADE_BULK=0 run "cp.async default"
for b in 1 2 3; do for b0 in 1 2 3; do PTB_SHADE_BLOCKS=$b PTB_SHADE_BLOCKS0=$b0 run "bulk blocks=$b blocks0=$b0"; done; done
