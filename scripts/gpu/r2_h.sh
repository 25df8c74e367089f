#!/bin/bash
# round-2 GPU call H: mean out of the contraction epilogue -- GP tests, bench A/B, fit probe
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -p no:cacheprovider -k "gp or precision or fused or plugin" > gpurun_out/r2h_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_tests.log
grep -n "passed\|failed\|^FAILED\|auto:" gpurun_out/r2h_tests.log | head -20
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-sort-hv > gpurun_out/r2h_bench.log 2>&1
DMO_GP_MEAN_SPLIT=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-sort-hv --no-cpu-baseline > gpurun_out/r2h_bench_split.log 2>&1
for f in gpurun_out/r2h_bench.log gpurun_out/r2h_bench_split.log; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['config']['gp_auto'], {k:round(v*d['ms_per_step'],3) for k,v in d['kernel_share_of_step'].items()}, 'frac', round(d['roofline']['frac'],4), d['clocks'])
PY
done
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
timeout 300 python /tmp/fit_probe.py > gpurun_out/r2h_fit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_fit_launches.csv python /tmp/fit_probe.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2h_fit_launches.csv > gpurun_out/r2h_fit_launches_summary.txt 2>&1
tail -3 gpurun_out/r2h_fit.log; head -12 gpurun_out/r2h_fit_launches_summary.txt
