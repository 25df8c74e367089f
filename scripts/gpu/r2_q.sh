#!/bin/bash
# round-2 GPU call Q: SMPSO swarms updated concurrently from worker contexts (one thread + stream per swarm)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_reference_loop.py -q -p no:cacheprovider -k "smpso or plugins_golden or c4 or epoch" > gpurun_out/r2q_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_tests.log
grep -n "passed\|failed\|^FAILED\|Error" gpurun_out/r2q_tests.log | head
DMO_VERBOSE_GENS=1 timeout 600 python scripts/config_sweep.py C4 > gpurun_out/r2q_c4_threads.log 2>&1
grep "ms/generation\|generation [0-9]" gpurun_out/r2q_c4_threads.log
DMOSOPT_B200_SMPSO_THREADS=0 timeout 600 python scripts/config_sweep.py C4 > gpurun_out/r2q_c4_serial.log 2>&1
grep "ms/generation" gpurun_out/r2q_c4_serial.log
