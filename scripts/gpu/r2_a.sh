#!/bin/bash
# round-2 GPU call A: new GP path (v3 + auto), reference-loop tests, bench A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "gp or duplicates or fused or egp or precision" > gpurun_out/r2a_gp_tests.log 2>&1
echo "gp tests rc=$?" >> gpurun_out/r2a_gp_tests.log
timeout 600 python -m pytest tests/test_gpu_reference_loop.py -x -q -s > gpurun_out/r2a_ref_tests.log 2>&1
echo "ref tests rc=$?" >> gpurun_out/r2a_ref_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_auto.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --precision tensor > gpurun_out/r2a_bench_tensor.log 2>&1
DMO_GP_NO_OVERLAP=1 timeout 300 python bench.py --no-cpu-baseline --precision tensor > gpurun_out/r2a_bench_tensor_nooverlap.log 2>&1
DMO_GP_TC=2 timeout 300 python bench.py --no-cpu-baseline --precision tensor > gpurun_out/r2a_bench_tensor_v2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_var_tc3 -s 3 -c 1 -o gpurun_out/r2a_gp_var_tc3 python bench.py --no-cpu-baseline --precision tensor --steps 2 --warmup 1 --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2a_ncu.log 2>&1
DMO_GP_NO_OVERLAP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_var_tc3 -s 3 -c 1 -o gpurun_out/r2a_gp_var_tc3_nooverlap python bench.py --no-cpu-baseline --precision tensor --steps 2 --warmup 1 --e2e-steps 1 --e2e-warmup 0 >> gpurun_out/r2a_ncu.log 2>&1
tail -5 gpurun_out/r2a_gp_tests.log gpurun_out/r2a_ref_tests.log
for f in gpurun_out/r2a_bench_*.log; do echo "== $f"; tail -1 $f | cut -c1-1500; done
