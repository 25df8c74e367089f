#!/bin/bash
# round-2 GPU call E (2 GPUs): N1 / TRS / benchmark tests, smoke, fit timing, 2-GPU bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -s -k "gp_fit or gpr_plugin or trs or benchmark_functions or hv3_tree or plugin" -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1
python - > gpurun_out/r2e_fit.log 2>&1 <<'PY'
import time, numpy as np, bench
import dmosopt_b200 as b2
from dmosopt_b200 import _lib as L
w = bench.workload(1024, 30, 3, 4096)
for fit in ("gpu", "gpu", "sklearn"):
    t0 = time.time(); sm = b2.GPR_Matern(w["Xtr"], w["Ytr"], 30, 3, w["xlb"], w["xub"], optimizer=None, fit=fit); print(fit, "fit+upload s", time.time() - t0, flush=True)
x = (w["Xtr"] - w["xlb"]) / (w["xub"] - w["xlb"])
yn = ((w["Ytr"] - w["Ytr"].mean(0)) / w["Ytr"].std(0)).T.copy()
L.profile_enable(True)
for _ in range(3):
    t0 = time.time(); L.gp_fit(x, yn, [1.0]*3, [np.full(30, 0.5)]*3, [1e-6]*3, want_L=False, want_alpha=False); print("lml-only trial, 3 objectives, s", time.time() - t0, flush=True)
print(L.profile_report())
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_2gpu.log 2>&1
tail -12 gpurun_out/r2e_tests.log; tail -3 gpurun_out/r2e_smoke.log; cat gpurun_out/r2e_fit.log | tail -12
tail -2 gpurun_out/r2e_bench_2gpu.log | cut -c1-3000
