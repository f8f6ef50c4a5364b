#!/bin/bash
# bench.py under torchrun at N GPUs (default 2 4 8) -> gpurun_out/<tag>_bench_n*.json; usage: gpurun --gpus 8 -- bash scripts/scale_run.sh tag "2 4 8"
tag=${1:-scale}
mkdir -p gpurun_out
for n in ${2:-2 4 8}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 \
    > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_n$n.json')); print('N=$n', 'Mpaths/s %.0f' % d['value'], 'ms %.2f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'trace frac %.3f' % d['roofline']['frac'], d['exchange'])" || tail -5 gpurun_out/${tag}_bench_n$n.err
done
