#!/bin/bash
# round-2 GPU call AC: truncation ranks by front peeling: parity against the chain, timings on the three objective sets, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "front_peeling or remove_worst or sortmo or nsga2_plugin or plugins_golden" > gpurun_out/r2ac_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2ac_tests.log
grep -n "passed\|failed\|^FAILED\|Error" gpurun_out/r2ac_tests.log | head; tail -30 gpurun_out/r2ac_tests.log | grep -n "assert\|^E " | head -20
for mode in 14 0; do
DMO_RANK_PEEL=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ac_bench_peel$mode.log 2>&1
tail -1 gpurun_out/r2ac_bench_peel$mode.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('peel=$mode', j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()}); print({k:(round(v['rank_truncate_ms'],3), round(v['hv_ms'],3)) for k,v in j['sort_hv'].items()})"
done
