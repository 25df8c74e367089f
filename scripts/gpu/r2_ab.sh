#!/bin/bash
# round-2 GPU call AB: rank chain with warp-leader polling: parity, timings (uniform / bench sets / other M), bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -p no:cacheprovider -k "rank or dda or sortmo or remove_worst or nsga2 or config or c2 or c3 or c4 or c5 or smpso or plugins_golden" > gpurun_out/r2ab_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2ab_tests.log
grep -n "passed\|failed\|^FAILED" gpurun_out/r2ab_tests.log | head
python scripts/rank_trace.py 131072 3 > gpurun_out/r2ab_trace_uniform.log 2>&1; head -13 gpurun_out/r2ab_trace_uniform.log
python scripts/rank_real.py 2 > gpurun_out/r2ab_rank_real.log 2>&1; grep "seg\|lex" gpurun_out/r2ab_rank_real.log
timeout 300 python scripts/kernel_sweep.py rank > gpurun_out/r2ab_kernel_sweep.log 2>&1; tail -12 gpurun_out/r2ab_kernel_sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sort-hv > gpurun_out/r2ab_bench.log 2>&1
tail -1 gpurun_out/r2ab_bench.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()})"
timeout 600 python scripts/config_sweep.py C4 C5 > gpurun_out/r2ab_sweep.log 2>&1; grep "ms/generation" gpurun_out/r2ab_sweep.log
