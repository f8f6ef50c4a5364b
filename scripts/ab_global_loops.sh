#!/bin/bash
# A/B of the global-memory traversal loops: the shipped library against lib/variant_*.so builds (see the Makefile's
# -DPTB_GLOBAL_RENDER_LOOP / -DPTB_GLOBAL_BATCH_LOOP).  usage: gpurun -- bash scripts/ab_global_loops.sh
L=path_tracer_ocaml_b200/lib
cp $L/libptb200.so /tmp/shipped.so
run() {
  echo "== $1"
  for mode in incoherent coherent; do timeout 600 python scripts/soup_probe.py 1000000 $mode soup 2>&1 | tail -1; done
  timeout 600 python scripts/soup_probe.py 1000000 incoherent mesh 2>&1 | tail -1
  timeout 600 python scripts/soup_probe.py 100000 incoherent soup 2>&1 | tail -1
  timeout 600 python scripts/soup_probe.py 10000 incoherent soup 2>&1 | tail -1
  timeout 600 python scripts/c3_probe.py 1000000 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 1M: trace %.3f ms, %.0f Mrays/s' % (d['ms_trace'], d['mrays_per_s']))"
  timeout 600 python scripts/c3_probe.py 10000000 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 10M: trace %.3f ms, %.0f Mrays/s' % (d['ms_trace'], d['mrays_per_s']))"
}
run shipped
for v in $L/variant_*.so; do cp $v $L/libptb200.so; run $(basename $v); done
cp /tmp/shipped.so $L/libptb200.so
