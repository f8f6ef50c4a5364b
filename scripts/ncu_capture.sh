#!/bin/bash
# development loop on one GPU: parity tests, a short bench (spp 64), then one full ncu capture of the first four
# k_trace / k_shade launches of an 8 Mi-path batch -> gpurun_out/exp_trace.ncu-rep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('Mpaths/s', d['value'], 'ms', d['ms_per_step'], 'trace ms', d['roofline']['trace_ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
[ "$1" = "--no-ncu" ] && exit 0
PTB_BATCH=8388608 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
PTB_BATCH=8388608 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 0 -c 4 \
   -f -o gpurun_out/exp_trace python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_exp.log 2>&1
echo done
