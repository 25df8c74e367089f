#!/bin/bash
# round-2 GPU call W (N GPUs): bench at --gpus $1 through torchrun, as the driver launches it
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2w_bench_${N}gpu.log 2>&1
echo "rc=$?"
grep '"metric"' gpurun_out/r2w_bench_${N}gpu.log | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'] if j.get('roofline') else None, {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()})"
tail -3 gpurun_out/r2w_bench_${N}gpu.log | cut -c1-300
