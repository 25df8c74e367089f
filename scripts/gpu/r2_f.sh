#!/bin/bash
# round-2 GPU call F (8 GPUs): strong-scaling bench lines at 8 and 4 GPUs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-sort-hv > gpurun_out/r2f_bench_${n}gpu.log 2>&1
  tail -1 gpurun_out/r2f_bench_${n}gpu.log | cut -c1-1200
done
