#!/bin/bash
# bytes per ray of the traversal kernel on the triangle configs (ncu, a handful of metrics; kernel times under ncu are
# not bench values): C3 mesh renders (the second render's bounce-0 and bounce-1 k_trace launches) and the C5 soup.
# usage: gpurun -- bash scripts/roofline_tri.sh
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for f in 1000000 10000000; do
  python scripts/c3_probe.py $f 4 > gpurun_out/r2_c3_${f}_plain.json 2> gpurun_out/r2_c3_${f}_plain.err; cat gpurun_out/r2_c3_${f}_plain.json
  # second render = launches 8.. (8 k_trace per render, one batch): take bounce 0 and 1 of it
  ncu --metrics $M --clock-control none -k regex:'k_trace' -s 8 -c 2 --csv --log-file gpurun_out/r2_c3_${f}_ncu.csv python scripts/c3_probe.py $f 4 > /dev/null 2>&1
done
for mode in incoherent coherent; do
  python scripts/soup_probe.py 1000000 $mode soup | tail -1 > gpurun_out/r2_soup_${mode}_plain.log; cat gpurun_out/r2_soup_${mode}_plain.log
  ncu --metrics $M --clock-control none -k regex:'k_trace' -s 2 -c 1 --csv --log-file gpurun_out/r2_soup_${mode}_ncu.csv python scripts/soup_probe.py 1000000 $mode soup > /dev/null 2>&1
done
python scripts/soup_probe.py 1000000 incoherent mesh | tail -1 > gpurun_out/r2_meshrays_plain.log; cat gpurun_out/r2_meshrays_plain.log
ncu --metrics $M --clock-control none -k regex:'k_trace' -s 2 -c 1 --csv --log-file gpurun_out/r2_meshrays_ncu.csv python scripts/soup_probe.py 1000000 incoherent mesh > /dev/null 2>&1
ls gpurun_out/r2_*ncu.csv
