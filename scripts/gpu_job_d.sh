#!/bin/bash
mkdir -p gpurun_out
for v in v4 noasm cur; do
  L=$PWD/path_tracer_ocaml_b200/lib/libptb200_$v.so; [ $v = cur ] && L=""
  PTB_LIB=$L python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d_$v.json 2>gpurun_out/d_$v.err
  python - <<PY
import json
try:
  d=json.load(open('gpurun_out/d_$v.json'))
  print('$v', 'Mpaths/s %.0f ms/step %.1f trace_ms %.1f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['trace_ms_per_step'],d['roofline']['frac']))
except Exception as e: print('$v','ERR',e)
PY
done
