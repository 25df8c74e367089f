#!/bin/bash
# round-2 GPU call Z: mean-only kernel evaluating ceil(d / 4) <= 4 dimension groups when d <= 16; C3 sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k "mean_only or gp_predict_golden or auto_on_a_fitted" > gpurun_out/r2z_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2z_tests.log
grep -n "passed\|failed\|^FAILED\|mean-only N" gpurun_out/r2z_tests.log | head -20
timeout 600 python scripts/config_sweep.py C3 > gpurun_out/r2z_c3.log 2>&1; grep "ms/generation" gpurun_out/r2z_c3.log
