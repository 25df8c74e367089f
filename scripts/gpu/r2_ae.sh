#!/bin/bash
# round-2 GPU call AE: peel grid resolution sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for gb in 9 10; do
DMO_PEEL_GBITS=$gb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ae_bench_g$gb.log 2>&1
tail -1 gpurun_out/r2ae_bench_g$gb.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('gbits=$gb', j['value'], j['ms_per_step'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()}); print({k:(round(v['rank_truncate_ms'],3), round(v['hv_ms'],3)) for k,v in j['sort_hv'].items()})"
done
