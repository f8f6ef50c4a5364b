#!/bin/bash
# A/B of lib/variant_*.so builds against the shipped library on the intersect_batch triangle workloads only
# (the render loops are not touched by these variants).  usage: gpurun -- bash scripts/ab_batch_variants.sh
L=path_tracer_ocaml_b200/lib
cp $L/libptb200.so /tmp/shipped.so
run() {
  echo "== $1"
  for mode in incoherent coherent; do timeout 600 python scripts/soup_probe.py 1000000 $mode soup 2>&1 | tail -1; done
  timeout 600 python scripts/soup_probe.py 100000 incoherent soup 2>&1 | tail -1
  timeout 600 python scripts/soup_probe.py 10000 incoherent soup 2>&1 | tail -1
}
run shipped
for v in $L/variant_*.so; do cp $v $L/libptb200.so; run $(basename $v); done
cp /tmp/shipped.so $L/libptb200.so
