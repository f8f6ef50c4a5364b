#!/bin/bash
# ncu --set full of the traversal kernel on the C5 triangle soup (third launch = warm).  usage: gpurun -- bash scripts/ncu_soup.sh tag [m] [incoherent|coherent] [soup|mesh]
tag=${1:-soup}; shift
mkdir -p gpurun_out
python scripts/soup_probe.py "$@" > gpurun_out/${tag}_plain.log 2>&1; tail -4 gpurun_out/${tag}_plain.log
ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 2 -c 1 -f -o gpurun_out/${tag} python scripts/soup_probe.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep
