#!/bin/bash
# round-2 GPU call G: compute-sanitizer on the round-2 kernels (small sizes), then the whole GPU suite, bench, smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/r2g_sanitizer.txt
: > $S
run_san() { echo "### compute-sanitizer --tool $1 :: $2" >> $S; timeout 900 compute-sanitizer --tool $1 --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "$2" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard|error" | head -12 >> $S; }
run_san memcheck "gp_fit_vs_scipy and (50 or 64) or hv3_tree and (130 or 1025) or benchmark_functions or trs_plugin or two_set or smpso_resident or plugins_golden or variation_loop"
run_san racecheck "gp_fit_vs_scipy and (50 or 64) or hv3_tree and (130 or 1025) or smpso_resident"
run_san synccheck "gp_fit_vs_scipy and (50 or 64) or hv3_tree and 1025 or gp_predict_tensor_path and 300"
cat $S
timeout 2400 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_tests.log
grep -n "passed\|failed\|^FAILED\|^ERROR" gpurun_out/r2g_tests.log | head -30
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench.log 2>&1
timeout 600 python bench.py > gpurun_out/r2g_bench_default.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
tail -2 gpurun_out/r2g_smoke.log
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
timeout 300 python /tmp/fit_probe.py > gpurun_out/r2g_fit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_fit_launches.csv python /tmp/fit_probe.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2g_fit_launches.csv > gpurun_out/r2g_fit_launches_summary.txt 2>&1
cat gpurun_out/r2g_fit.log | tail -3; head -12 gpurun_out/r2g_fit_launches_summary.txt
timeout 600 python scripts/config_sweep.py C2 C3 C4 C5 > gpurun_out/r2g_config_sweep.log 2>&1
cat gpurun_out/r2g_config_sweep.log | tail -8
for f in gpurun_out/r2g_bench.log gpurun_out/r2g_bench_default.log; do tail -1 $f | cut -c1-900; done
