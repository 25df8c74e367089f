#!/bin/bash
# round-2 GPU call V: exact hypervolume for 6 .. 8 objectives (hv_many.cu): golden + oracle + cross-check, timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "test_hv" --durations=8 > gpurun_out/r2v_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2v_tests.log
grep -n "passed\|failed\|^FAILED\|Error\|assert" gpurun_out/r2v_tests.log | head -20; grep -A10 "slowest" gpurun_out/r2v_tests.log | head -12
python - > gpurun_out/r2v_timing.log 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from dmosopt_b200 import _lib as L
rng = np.random.default_rng(0)
for n, M in ((100, 6), (300, 6), (600, 6), (200, 7), (400, 7), (150, 8), (300, 8)):
    x = rng.random((n, M)); P = x / np.linalg.norm(x, axis=1, keepdims=True) * (1 + 0.1 * rng.random((n, 1)))
    ref = np.full(M, 1.2)
    L.hypervolume(P[:20], ref)
    t0 = time.perf_counter(); v = L.hypervolume(P, ref); dt = time.perf_counter() - t0
    print(f"n={n} M={M}: hv={v:.10g} in {dt*1e3:.1f} ms", flush=True)
PY
cat gpurun_out/r2v_timing.log
