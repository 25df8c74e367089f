#!/bin/bash
# round-2 GPU call P: fused K_* + mean kernel after the flush rewrite: parity subset, bench fused / unfused
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "tensor_path or benchmarked_shape or golden_cases or fitted_model" > gpurun_out/r2p_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_tests.log
grep -n "passed\|failed\|^FAILED" gpurun_out/r2p_tests.log | head
for mode in 1 0; do
DMO_GP_FUSED=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sort-hv > gpurun_out/r2p_bench_fused$mode.log 2>&1
tail -1 gpurun_out/r2p_bench_fused$mode.log | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fused=$mode', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['gpu_launches'], {k: round(v*j['ms_per_step'],3) for k,v in j['kernel_share_of_step'].items()})"
done
