#!/bin/bash
# round-2 GPU call AG: compute-sanitizer on the peel / many-objective HV kernels; ncu --set full of peel_flag_kernel and the final kstar_mean_kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/r2ag_sanitizer.txt
: > $S
run_san() { echo "### compute-sanitizer --tool $1 :: $2" >> $S; timeout 900 compute-sanitizer --tool $1 --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "$2" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|Race|hazard|error" | head -12 >> $S; }
run_san memcheck "front_peeling and 20000 and (layers or ties or duplicates) or six_to_eight or limit_set_recursion and 60"
run_san racecheck "front_peeling and 20000 and layers or limit_set_recursion and 60"
cat $S
B="python bench.py --no-cpu-baseline --no-sort-hv --steps 2 --warmup 1 --e2e-steps 1 --e2e-warmup 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:peel_flag_kernel -s 10 -c 1 -o gpurun_out/r2ag_peel_flag $B > gpurun_out/r2ag_1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kstar_mean_kernel -s 2 -c 1 -o gpurun_out/r2ag_kstar_mean $B > gpurun_out/r2ag_2.log 2>&1
ls -la gpurun_out/r2ag_*.ncu-rep
