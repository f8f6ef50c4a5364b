#!/bin/bash
# one GPU box: parity tests, smoke, a short bench.  Usage: gpurun -- bash scripts/gpu_check.sh [tag]
tag=${1:-check}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -25 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
