#!/bin/bash
# round-2 GPU call B: diagnose the overlapped K* / variance pipeline
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r2b_probe.log
: > $LOG
run() { echo "--- $*" >> $LOG; ( env "$@" timeout 120 python scripts/gpu/gp_overlap_probe.py $P >> $LOG 2>&1 ); echo "rc=$?" >> $LOG; }
P=4608
run DMO_GP_NO_OVERLAP=1
run DMO_GP_DBG=4
run DMO_GP_DBG=12
run DMO_GP_DBG=0
run DMO_GP_DBG=1
run DMO_GP_DBG=2
run DMO_GP_DBG=3
run CUDA_LAUNCH_BLOCKING=1
P=65536
run DMO_GP_NO_OVERLAP=1
run DMO_GP_DBG=4
run DMO_GP_DBG=0
echo "--- sanitizer" >> $LOG
P=4608
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu/gp_overlap_probe.py 4608 > gpurun_out/r2b_sanitizer.log 2>&1
tail -40 gpurun_out/r2b_sanitizer.log >> $LOG
grep -v "^Traceback\|^  File\|^    " $LOG | tail -60
# --- everything else with the in-line (non-overlapped) pipeline, so that one open bug does not hide the rest
export DMO_GP_NO_OVERLAP=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "gp or hv or precision or fused" > gpurun_out/r2b_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_tests.log
timeout 600 python -m pytest tests/test_gpu_reference_loop.py -q -s > gpurun_out/r2b_ref_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_ref_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.log 2>&1
tail -15 gpurun_out/r2b_tests.log; tail -15 gpurun_out/r2b_ref_tests.log; tail -2 gpurun_out/r2b_bench.log | cut -c1-3000
