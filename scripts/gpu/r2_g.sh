#!/bin/bash
# round-2 GPU call G: compute-sanitizer on the round-2 kernels (small sizes), then the whole GPU suite, bench, smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/r2g_sanitizer.txt
: > $S
run_san() { echo "### compute-sanitizer --tool $1 :: $2" >> $S; timeout 900 compute-sanitizer --tool $1 --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "$2" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard|error" | head -12 >> $S; }
run_san memcheck "gp_fit_vs_scipy and (50 or 64) or hv3_tree and (130 or 1025) or benchmark_functions or trs_plugin or two_set or smpso_resident or plugins_golden or variation_loop"
run_san racecheck "gp_fit_vs_scipy and (50 or 64) or hv3_tree and (130 or 1025) or smpso_resident"
run_san synccheck "gp_fit_vs_scipy and (50 or 64) or hv3_tree and 1025 or gp_predict_tensor_path and 300"
cat $S
timeout 2400 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_tests.log
grep -n "passed\|failed\|^FAILED\|^ERROR" gpurun_out/r2g_tests.log | head -30
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench.log 2>&1
timeout 600 python bench.py > gpurun_out/r2g_bench_default.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
tail -2 gpurun_out/r2g_smoke.log
timeout 600 python scripts/config_sweep.py C2 C3 C4 C5 > gpurun_out/r2g_config_sweep.log 2>&1
cat gpurun_out/r2g_config_sweep.log | tail -8
for f in gpurun_out/r2g_bench.log gpurun_out/r2g_bench_default.log; do tail -1 $f | cut -c1-900; done
