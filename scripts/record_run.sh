#!/bin/bash
# round-1 record run (one GPU): parity tests, smoke, headline bench + reference arm, ncu launch list and full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log; tail -2 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench.err
python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
PTB_BATCH=8388608 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
PTB_BATCH=8388608 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 0 -c 4 \
   -f -o gpurun_out/final_trace python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python tests/configs_bench.py --skip-sweep > gpurun_out/final_configs_c123.json 2> gpurun_out/final_configs.err
echo done
