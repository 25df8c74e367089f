#!/bin/bash
# round-2 GPU call AJ: rank chain with the second barrier per tile: racecheck (segmented + lexicographic), parity, timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/r2aj_sanitizer.txt
: > $S
echo "### compute-sanitizer --tool racecheck :: rank chain, segmented order (n = 6000, M = 3 and 2) and lexicographic order (M = 4), peeling off" >> $S
DMO_RANK_PEEL=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python -c "
import numpy as np
from dmosopt_b200 import _lib as L
rng=np.random.default_rng(0)
for M in (3, 2, 4):
    Y=rng.random((6000,M)); r=L.rank_nd(Y); print('M', M, 'fronts', r.max()+1)
" 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|fronts" | head -8 >> $S
cat $S
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "rank or front_peeling or remove_worst or sortmo" > gpurun_out/r2aj_tests.log 2>&1; tail -2 gpurun_out/r2aj_tests.log
timeout 300 python scripts/kernel_sweep.py rank > gpurun_out/r2aj_kernel_sweep.log 2>&1; tail -6 gpurun_out/r2aj_kernel_sweep.log
