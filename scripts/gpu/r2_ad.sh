#!/bin/bash
# round-2 GPU call AD: grid cells follow the number of distinct ids; peel + rank-0 filter tests, HV tests, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider --durations=6 -k "front_peeling or test_hv or remove_worst or sortmo or nsga2_plugin or rank" > gpurun_out/r2ad_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2ad_tests.log
grep -n "passed\|failed\|^FAILED\|Error" gpurun_out/r2ad_tests.log | head; grep -A8 "slowest" gpurun_out/r2ad_tests.log | head -10
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ad_bench.log 2>&1
tail -1 gpurun_out/r2ad_bench.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['value'], j['ms_per_step'], j['e2e']['value'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()}); print({k:(round(v['rank_truncate_ms'],3), round(v['hv_ms'],3)) for k,v in j['sort_hv'].items()})"
