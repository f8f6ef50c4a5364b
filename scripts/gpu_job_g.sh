#!/bin/bash
# round-1 record run: parity tests, headline bench (N=1) + reference arm, ncu launch list of a short run
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_r1_v6.json 2> gpurun_out/bench_r1_v6.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_r1_v6_reference.json 2>> gpurun_out/bench_r1_v6.err
python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo done
