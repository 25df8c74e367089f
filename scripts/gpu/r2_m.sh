#!/bin/bash
# round-2 GPU call M: mean-only direct kernel (K_* never written): parity, timing against the K_* + alpha-pass route, config sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k "mean_only or benchmarked_shape or golden_cases or fitted_model or tensor_path or linear_mean" > gpurun_out/r2m_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_tests.log
grep -n "passed\|failed\|^FAILED\|mean-only N" gpurun_out/r2m_tests.log | head -30
cat > /tmp/mean_probe.py <<'PY'
import sys, time, os, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from dmosopt_b200 import _lib as L
import dmosopt_b200 as b2
for (P, d, M, N) in ((65536, 30, 3, 4096), (65536, 12, 3, 4096), (327680, 22, 5, 4096)):
    w = bench.workload(P, d, M, N)
    sm = b2.GPR_Matern(w["Xtr"], w["Ytr"], d, M, w["xlb"], w["xub"], optimizer=None)
    X = np.random.default_rng(1).random((P, d))
    for mode in ("1", "0"):
        os.environ["DMO_GP_MEAN_DIRECT"] = mode
        sm.evaluate(X[:1024])
        for _ in range(2): sm.evaluate(X)
        L.synchronize(); L.profile_enable(True)
        t0 = time.perf_counter()
        for _ in range(5): y = sm.evaluate(X)
        L.synchronize(); dt = (time.perf_counter() - t0) / 5
        rep = {k: (round(v[0] / 5, 3), v[1] // 5) for k, v in L.profile_report().items()}
        L.profile_enable(False)
        print(f"P={P} d={d} M={M} N={N} direct={mode}: evaluate {dt*1e3:.2f} ms (host arrays in/out); kernels ms/launches per call: {rep}; info {sm._gp.auto_info()}", flush=True)
PY
DMO_GP_VERBOSE=1 timeout 600 python /tmp/mean_probe.py > gpurun_out/r2m_mean_probe.log 2>&1
cat gpurun_out/r2m_mean_probe.log | cut -c1-600
DMO_VERBOSE_GENS=0 timeout 900 python scripts/config_sweep.py > gpurun_out/r2m_config_sweep.log 2>&1
grep "ms/generation" gpurun_out/r2m_config_sweep.log
