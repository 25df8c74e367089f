#!/bin/bash
# racecheck of the peel and many-objective HV kernels after the rank-chain barrier fix
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/r2ak_sanitizer.txt
: > $S
echo "### compute-sanitizer --tool racecheck :: front_peeling and 20000 and layers or limit_set_recursion and 60" >> $S
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "front_peeling and 20000 and layers or limit_set_recursion and 60" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Race reported" | head -8 >> $S
cat $S
