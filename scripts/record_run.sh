#!/bin/bash
# The round's record on one GPU: parity tests, smoke, headline bench + reference arm, ncu launch list, one full capture
# of the first k_trace / k_shade launches, k_shade's DRAM bytes at a production-size batch, the other configs.
# usage: gpurun --timeout 2400 -- bash scripts/record_run.sh [tag]   -> gpurun_out/<tag>_* (copied into profiles/ by hand)
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_pytest_gpu.log
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench.err
python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2>> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
# launch list of a bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# full capture: bounce-0 and bounce-1 launches of an 8 Mi-path batch
PTB_BATCH=8388608 python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
PTB_BATCH=8388608 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 0 -c 4 \
   -f -o gpurun_out/${tag}_trace python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
# k_shade DRAM bytes per hit at a 64 Mi-path batch (single-pass metrics: no kernel replay)
PTB_BATCH=67108864 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
   -k regex:'k_shade' -s 0 -c 8 --csv --log-file gpurun_out/${tag}_shade_bytes.csv python bench.py --spp 8 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_shade.log 2>&1
PTB_BATCH=67108864 PTB_TIMING=1 python bench.py --spp 8 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_shade_bytes_plain.log 2>&1
python tests/configs_bench.py --skip-sweep > gpurun_out/${tag}_configs_c123.json 2> gpurun_out/${tag}_configs.err
echo done
