#!/bin/bash
# racecheck of the segmented rank chain (experiment)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DMO_RANK_PEEL=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python -c "
import numpy as np
from dmosopt_b200 import _lib as L
rng=np.random.default_rng(0); Y=rng.random((6000,3))
r=L.rank_nd(Y); print('fronts', r.max()+1)
" 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|fronts|Read access|Write access" | head -12
