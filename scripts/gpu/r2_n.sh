#!/bin/bash
# round-2 GPU call N: K_* producer fused with the mean (kstar_mean_kernel): parity, kernel times, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k "gp_ or egp" > gpurun_out/r2n_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_tests.log
grep -n "passed\|failed\|^FAILED\|tensor path N\|tensor, N=" gpurun_out/r2n_tests.log | head -30
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sort-hv > gpurun_out/r2n_bench_fused.log 2>&1
tail -1 gpurun_out/r2n_bench_fused.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fused  ', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['gpu_launches'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()})"
DMO_GP_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sort-hv > gpurun_out/r2n_bench_unfused.log 2>&1
tail -1 gpurun_out/r2n_bench_unfused.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('unfused', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['gpu_launches'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()})"
