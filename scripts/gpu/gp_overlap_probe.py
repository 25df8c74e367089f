"""One tensor-path predict at the BASELINE model size (N = 4096, d = 30, M = 3) for the DMO_GP_* diagnostics switches.
usage: python scripts/gpu/gp_overlap_probe.py P   (environment: DMO_GP_NO_OVERLAP, DMO_GP_DBG, DMO_GP_TC)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmosopt_b200 import _lib as L  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4608
N, d, M = 4096, 30, 3
w = bench.workload(P, d, M, N)
cache = "/tmp/dmo_probe_state.npz"  # not under gpurun_out/: it is 400 MB and gpurun copies that directory back
if os.path.exists(cache):
    z = np.load(cache)
    alpha, Lf, ym, ys = z["alpha"], z["L"], z["ym"], z["ys"]
else:
    from oracle import gp

    st = gp.fit_fixed(w["Xtr"], w["Ytr"], w["xlb"], w["xub"], 1.0, 0.5, 1e-6)
    alpha = np.stack([o.alpha for o in st.objectives])
    Lf = np.stack([o.L for o in st.objectives])
    ym = np.array([o.y_mean for o in st.objectives])
    ys = np.array([o.y_std for o in st.objectives])
    np.savez(cache, alpha=alpha, L=Lf, ym=ym, ys=ys)
h = L.GPHandle(w["Xtr"], alpha, Lf, [1.0] * M, [np.full(d, 0.5)] * M, [1e-6] * M, ym, ys, w["xlb"], w["xub"])
X = np.random.default_rng(1).random((P, d))
m64, v64 = h.predict(X[:512], precision=L.GP_FP64)
for rep in range(3):
    t0 = time.perf_counter()
    mt, vt = h.predict(X, precision=L.GP_TENSOR)
    dt = time.perf_counter() - t0
prior = (1.0 + 1e-6) * ys**2
print(f"OK P={P} env={ {k: v for k, v in os.environ.items() if k.startswith('DMO_GP')} } {dt*1e3:.2f} ms  var err/prior {np.max(np.abs(vt[:512]-v64)/prior):.2e} mean err {np.max(np.abs(mt[:512]-m64)/np.maximum(np.abs(m64), ys)):.2e}", flush=True)
