#!/bin/bash
# round-2 GPU call AF (final): whole GPU suite, smoke, bench (driver flags + default), reference arm, crowding variant, launch list, config sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/r2af_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2af_tests.log
grep -n "passed\|failed\|^FAILED\|^ERROR" gpurun_out/r2af_tests.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2af_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2af_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2af_bench.log 2>&1
timeout 600 python bench.py > gpurun_out/r2af_bench_default.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2af_bench_ref.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-sort-hv --distance-metric crowding --steps 20 --warmup 5 > gpurun_out/r2af_bench_crowding.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2af_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sort-hv --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2af_launches_run.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2af_launches.csv > gpurun_out/r2af_launches_summary.txt 2>&1
rm -f gpurun_out/r2af_launches.csv
timeout 900 python scripts/config_sweep.py > gpurun_out/r2af_config_sweep.log 2>&1; grep "ms/generation\|HV-improvement" gpurun_out/r2af_config_sweep.log
for f in gpurun_out/r2af_bench.log gpurun_out/r2af_bench_default.log gpurun_out/r2af_bench_ref.log gpurun_out/r2af_bench_crowding.log; do echo "== $f"; tail -1 $f | cut -c1-330; done
