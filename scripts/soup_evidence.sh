#!/bin/bash
# Evidence for the pre-split + sorted traversal of the C5 triangle soup: plain timings, one ncu --set full capture of the
# traversal kernel (incoherent rays) and the bytes-per-ray metrics of the incoherent and coherent launches.
# usage: gpurun -- bash scripts/soup_evidence.sh [tag]   -> gpurun_out/<tag>_*
tag=${1:-r2p_soup}
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for mode in incoherent coherent; do
  PTB_BVH_TIMING=1 python scripts/soup_probe.py 1000000 $mode soup > gpurun_out/${tag}_${mode}_plain.log 2>&1; tail -3 gpurun_out/${tag}_${mode}_plain.log
  ncu --metrics $M --clock-control none -k regex:'k_trace' -s 2 -c 1 --csv --log-file gpurun_out/${tag}_${mode}_bytes_ncu.csv python scripts/soup_probe.py 1000000 $mode soup > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 2 -c 1 -f -o gpurun_out/${tag} python scripts/soup_probe.py 1000000 incoherent soup > gpurun_out/${tag}_ncu.log 2>&1
# the sort's own kernels (times under ncu are serialised, cold-cache)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches.csv python scripts/soup_probe.py 1000000 incoherent soup > /dev/null 2>&1
ls -la gpurun_out/${tag}*
