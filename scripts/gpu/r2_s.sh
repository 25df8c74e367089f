#!/bin/bash
# round-2 GPU call S: duplicates keyed by a row projection; AGE-MOEA (C2) again
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_loop.py -q -p no:cacheprovider -k "duplicates or install or regress or plugins_golden or nsga2_plugin" > gpurun_out/r2s_tests.log 2>&1
tail -4 gpurun_out/r2s_tests.log
DMO_PROFILE=1 timeout 600 python scripts/config_sweep.py C2 > gpurun_out/r2s_c2_profile.log 2>&1
grep -n "ms/generation" gpurun_out/r2s_c2_profile.log; grep -A16 "cumulative" gpurun_out/r2s_c2_profile.log | cut -c1-150 | head -22
