#!/bin/bash
# round-2 GPU call U: compute-sanitizer on the kernels added after call G (fused K_* + mean, mean-only, resident MO-CMA-ES steps,
# projection-keyed duplicates, register-resident potrf)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/r2u_sanitizer.txt
: > $S
run_san() { echo "### compute-sanitizer --tool $1 :: $2" >> $S; timeout 1200 compute-sanitizer --tool $1 --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "$2" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|Race|hazard|error" | head -12 >> $S; }
run_san memcheck "gp_mean_only_kernel and (300 or 513 or 777) or gp_predict_tensor_path and (300 or 600) or cmaes_resident or test_duplicates or gp_fit_vs_scipy and (50 or 64) or test_cmaes_kernels"
run_san racecheck "gp_mean_only_kernel and (300 or 513) or gp_predict_tensor_path and 300 or cmaes_resident or gp_fit_vs_scipy and 64"
run_san synccheck "gp_mean_only_kernel and 300 or gp_predict_tensor_path and 300 or gp_fit_vs_scipy and 64 or cmaes_resident"
cat $S
