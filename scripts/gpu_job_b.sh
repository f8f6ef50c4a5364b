#!/bin/bash
# A/B: new build vs lib/libptb200_base.so; parity tests on the new build
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in base new; do
  L=""; [ $v = base ] && L=$PWD/path_tracer_ocaml_b200/lib/libptb200_base.so
  PTB_LIB=$L python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err
done
for r in 8 20; do
  PTB_REFILL=$r python bench.py --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_new_refill_$r.json 2>/dev/null
done
python scripts/dev_check.py > gpurun_out/dev_check3.log 2>&1
echo done
