#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DMO_RANK_PEEL=0 timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 12 python -c "
import numpy as np
from dmosopt_b200 import _lib as L
rng=np.random.default_rng(0); Y=rng.random((6000,3))
r=L.rank_nd(Y); print('fronts', r.max()+1)
" > gpurun_out/r2ai_hazards.log 2>&1
grep -E "hazard detected|Write Thread|Read Thread|Current Value|SUMMARY" gpurun_out/r2ai_hazards.log | head -40
