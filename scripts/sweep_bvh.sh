#!/bin/bash
# trace-kernel time against the tree builder's collapse (PTB_BVH_DP) and node cost (PTB_BVH_CT); usage: gpurun -- bash scripts/sweep_bvh.sh
mkdir -p gpurun_out
for dp in 0 1; do for ct in ${1:-1.2 2 3 4}; do
  PTB_BVH_DP=$dp PTB_BVH_CT=$ct python bench.py --spp 128 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dp=$dp ct=$ct', 'Mpaths/s %.0f' % d['value'], 'trace ms %.2f' % d['roofline']['trace_ms_per_step'], 'frac %.4f' % d['roofline']['frac'])"
done; done | tee gpurun_out/sweep_bvh.log
