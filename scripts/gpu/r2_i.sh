#!/bin/bash
# round-2 GPU call I: gp_fit rewrite -- tests, timing; calibration trace; config sweep C5
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "gp_fit or gpr_plugin or golden_fp64 or auto" > gpurun_out/r2i_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2i_tests.log
grep -n "passed\|failed\|^FAILED" gpurun_out/r2i_tests.log | head
cat > /tmp/fit_probe.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from dmosopt_b200 import _lib as L
import dmosopt_b200 as b2
w = bench.workload(1024, 30, 3, 4096)
x = (w["Xtr"] - w["xlb"]) / (w["xub"] - w["xlb"])
yn = ((w["Ytr"] - w["Ytr"].mean(0)) / w["Ytr"].std(0)).T.copy()
for _ in range(3):
    t0 = time.time(); L.gp_fit(x, yn[:1], [1.0], [np.full(30, 0.5)], [1e-6], want_L=False, want_alpha=False); print("lml-only, 1 objective, N=4096: s", time.time() - t0, flush=True)
for _ in range(2):
    t0 = time.time(); L.gp_fit(x, yn, [1.0]*3, [np.full(30, 0.5)]*3, [1e-6]*3); print("full fit (L, alpha, lml), 3 objectives: s", time.time() - t0, flush=True)
for _ in range(2):
    t0 = time.time(); sm = b2.GPR_Matern(w["Xtr"], w["Ytr"], 30, 3, w["xlb"], w["xub"], optimizer=None); print("GPR_Matern fit+upload s", time.time() - t0, flush=True)
sm.predict(w["X0"][:600])
PY
DMO_GP_VERBOSE=1 timeout 300 python /tmp/fit_probe.py > gpurun_out/r2i_fit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_fit_launches.csv python /tmp/fit_probe.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2i_fit_launches.csv > gpurun_out/r2i_fit_launches_summary.txt 2>&1
cat gpurun_out/r2i_fit.log | tail -9; head -14 gpurun_out/r2i_fit_launches_summary.txt
timeout 600 python scripts/config_sweep.py C5 > gpurun_out/r2i_config_sweep.log 2>&1; tail -2 gpurun_out/r2i_config_sweep.log
