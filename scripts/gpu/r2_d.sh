#!/bin/bash
# round-2 GPU call D: whole GPU suite (all failures reported), bench with driver flags
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu --durations=12 -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2d_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2d_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2d_smoke.log 2>&1
grep -n "passed\|failed\|^FAILED\|^ERROR\|Error" gpurun_out/r2d_tests.log | head -40
grep -n "ms/generation" gpurun_out/r2d_tests.log
tail -3 gpurun_out/r2d_smoke.log
tail -1 gpurun_out/r2d_bench.log | cut -c1-6000
