#!/bin/bash
# round-2 GPU call J: gp_fit (register-resident potrf) tests + timing, C5 with a host profile, C4
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "gp_fit or gpr_plugin" > gpurun_out/r2j_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2j_tests.log
grep -n "passed\|failed\|^FAILED" gpurun_out/r2j_tests.log | head
cat > /tmp/fit_probe.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from dmosopt_b200 import _lib as L
w = bench.workload(1024, 30, 3, 4096)
x = (w["Xtr"] - w["xlb"]) / (w["xub"] - w["xlb"])
yn = ((w["Ytr"] - w["Ytr"].mean(0)) / w["Ytr"].std(0)).T.copy()
for _ in range(3):
    t0 = time.time(); L.gp_fit(x, yn[:1], [1.0], [np.full(30, 0.5)], [1e-6], want_L=False, want_alpha=False); print("lml-only, 1 objective, N=4096: s", time.time() - t0, flush=True)
PY
timeout 300 python /tmp/fit_probe.py > gpurun_out/r2j_fit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_fit_launches.csv python /tmp/fit_probe.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2j_fit_launches.csv > gpurun_out/r2j_fit_launches_summary.txt 2>&1
tail -3 gpurun_out/r2j_fit.log; head -9 gpurun_out/r2j_fit_launches_summary.txt
DMO_PROFILE=1 timeout 600 python scripts/config_sweep.py C5 > gpurun_out/r2j_c5_profile.log 2>&1
grep -n "ms/generation" gpurun_out/r2j_c5_profile.log; grep -A34 "cumulative" gpurun_out/r2j_c5_profile.log | cut -c1-150 | head -42
timeout 600 python scripts/config_sweep.py C4 C5 > gpurun_out/r2j_config_sweep.log 2>&1; grep "ms/generation" gpurun_out/r2j_config_sweep.log
