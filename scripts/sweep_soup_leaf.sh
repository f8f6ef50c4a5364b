#!/bin/bash
# C5 soup (1M triangles, incoherent rays) against the host builder's leaf size and node cost.  usage: gpurun -- bash scripts/sweep_soup_leaf.sh
mkdir -p gpurun_out
for leaf in 1 2 4 8; do for ct in ${1:-1.2}; do
  echo "leaf=$leaf ct=$ct"; PTB_BUILDER=host PTB_BVH_LEAF=$leaf PTB_BVH_CT=$ct timeout 600 python scripts/soup_probe.py 1000000 incoherent soup 2>&1 | grep -E "^tree|Grays" | tail -2
done; done | tee gpurun_out/sweep_soup_leaf.log
