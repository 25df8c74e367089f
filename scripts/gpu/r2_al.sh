#!/bin/bash
# peel flag kernel with sixteen loads in flight: parity subset + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "front_peeling" > gpurun_out/r2al_tests.log 2>&1; tail -1 gpurun_out/r2al_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2al_bench.log 2>&1
tail -1 gpurun_out/r2al_bench.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['value'], j['ms_per_step'], j['e2e']['value'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()}); print({k:(round(v['rank_truncate_ms'],3), round(v['hv_ms'],3)) for k,v in j['sort_hv'].items()})"
