#!/bin/bash
# round-2 GPU call L: MO-CMA-ES with the parent ranking overlapped with the host RNG; per-generation times; all config tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DMO_VERBOSE_GENS=1 timeout 600 python scripts/config_sweep.py C5 > gpurun_out/r2l_c5.log 2>&1
cat gpurun_out/r2l_c5.log | tail -8
DMO_VERBOSE_GENS=1 timeout 900 python scripts/config_sweep.py > gpurun_out/r2l_config_sweep.log 2>&1
grep "ms/generation\|generation [0-9]" gpurun_out/r2l_config_sweep.log
timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_reference_loop.py -q -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_tests.log
tail -6 gpurun_out/r2l_tests.log
