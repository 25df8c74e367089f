#!/bin/bash
# round-2 GPU call K: device-resident MO-CMA-ES generation / update (positions and step sizes in HBM)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_loop.py -q -p no:cacheprovider -k "cmaes or plugins_golden or epoch or install" > gpurun_out/r2k_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2k_tests.log
grep -n "passed\|failed\|^FAILED\|Error" gpurun_out/r2k_tests.log | head -20
DMO_PROFILE=1 timeout 600 python scripts/config_sweep.py C5 > gpurun_out/r2k_c5_profile.log 2>&1
grep -n "ms/generation" gpurun_out/r2k_c5_profile.log; grep -A34 "cumulative" gpurun_out/r2k_c5_profile.log | cut -c1-150 | head -42
timeout 900 python -m pytest tests/test_gpu_configs.py -q -p no:cacheprovider -k "c5" > gpurun_out/r2k_c5_test.log 2>&1
echo "rc=$?" >> gpurun_out/r2k_c5_test.log
tail -5 gpurun_out/r2k_c5_test.log
