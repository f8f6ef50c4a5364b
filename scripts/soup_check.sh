#!/bin/bash
# the triangle workloads after a change of the global-memory traversal: GPU parity tests, then the C5 soup and the C3
# mesh as stand-alone timings.  usage: gpurun -- bash scripts/soup_check.sh [tag]
tag=${1:-soup}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
for mode in incoherent coherent; do
  timeout 600 python scripts/soup_probe.py 1000000 $mode soup 2>&1 | tail -2 | tee -a gpurun_out/${tag}_soup.log
done
timeout 600 python scripts/soup_probe.py 1000000 incoherent mesh 2>&1 | tail -1 | tee -a gpurun_out/${tag}_soup.log
timeout 600 python scripts/c3_probe.py 1000000 4 2>/dev/null | tee -a gpurun_out/${tag}_soup.log
timeout 600 python scripts/c3_probe.py 10000000 4 2>/dev/null | tee -a gpurun_out/${tag}_soup.log
