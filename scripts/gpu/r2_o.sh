#!/bin/bash
# round-2 GPU call O: ncu --set full of the fused K_* + mean kernel and of the mean-only kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kstar_mean_kernel -s 2 -c 1 -o gpurun_out/r2o_kstar_mean python bench.py --no-cpu-baseline --no-sort-hv --steps 2 --warmup 1 --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2o_ncu1.log 2>&1
tail -2 gpurun_out/r2o_ncu1.log | cut -c1-300
cat > /tmp/mo.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench, dmosopt_b200 as b2
w = bench.workload(65536, 30, 3, 4096)
sm = b2.GPR_Matern(w["Xtr"], w["Ytr"], 30, 3, w["xlb"], w["xub"], optimizer=None)
X = np.random.default_rng(1).random((65536, 30))
for _ in range(3): sm.evaluate(X)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_mean_direct_kernel -s 3 -c 1 -o gpurun_out/r2o_mean_direct python /tmp/mo.py > gpurun_out/r2o_ncu2.log 2>&1
tail -2 gpurun_out/r2o_ncu2.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
