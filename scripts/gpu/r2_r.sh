#!/bin/bash
# round-2 GPU call R: AGE-MOEA (C2) host profile after vectorising the later-front scores
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DMO_PROFILE=1 timeout 600 python scripts/config_sweep.py C2 > gpurun_out/r2r_c2_profile.log 2>&1
grep -n "ms/generation" gpurun_out/r2r_c2_profile.log; grep -A34 "cumulative" gpurun_out/r2r_c2_profile.log | cut -c1-150 | head -40
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -p no:cacheprovider -k "age or c2 or plugins_golden" > gpurun_out/r2r_tests.log 2>&1
tail -3 gpurun_out/r2r_tests.log
