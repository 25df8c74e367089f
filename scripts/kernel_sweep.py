"""GPU micro-timings of the individual hot-path kernels (CUDA events through the library's profile timers)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmosopt_b200 import _lib as L  # noqa: E402


def timed(fn, reps=3):
    fn()
    L.synchronize()
    best = 1e9
    for _ in range(reps):
        L.timer_begin()
        fn()
        best = min(best, L.timer_end())
    return best


def main():
    L.context()
    rng = np.random.default_rng(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "rank"
    if what == "rank":
        for n, M, kind in [(131072, 3, "uniform"), (131072, 3, "sphere"), (131072, 2, "uniform"), (16384, 2, "uniform"), (65536, 5, "uniform")]:
            Y = rng.random((n, M))
            if kind == "sphere":
                Y = Y / np.linalg.norm(Y, axis=1, keepdims=True) * (1 + 0.01 * rng.random((n, 1)))
            d = L.DeviceArray((n, M)).upload(Y)
            r = L.DeviceArray((n,), np.int32)
            lib, ctx = L.load_library(), L.context()
            L.profile_enable(True)
            ms = timed(lambda: L._check(lib.dmo_rank_nd(ctx, d.ptr, n, M, r.ptr), "rank"))
            rep = L.profile_report()
            L.profile_enable(False)
            rk = r.download()
            print(f"rank n={n} M={M} {kind}: total {ms:.3f} ms, chain {rep['rank_chain'][0] / rep['rank_chain'][1]:.3f} ms, fronts {rk.max() + 1}, OCC={os.environ.get('DMO_RANK_OCC', 'default')}", flush=True)


def gp_sweep():
    """GP posterior at the BASELINE shape: fp64 vs tensor path versions (DMO_GP_TC=1|2), accuracy and time."""
    L.context()
    rng = np.random.default_rng(1)
    N, d, M, P = 4096, 30, 3, 65536
    from oracle import gp as ogp

    Xtr = rng.random((N, d))
    Ytr = np.column_stack([np.sin(3 * Xtr[:, :4].sum(axis=1) + k) + Xtr[:, 4 + k] ** 2 for k in range(M)])
    st = ogp.fit_fixed(Xtr, Ytr, np.zeros(d), np.ones(d), 1.0, 0.5, 1e-6)
    h = L.GPHandle(st.X_train, np.stack([o.alpha for o in st.objectives]), np.stack([o.L for o in st.objectives]), [o.constant for o in st.objectives],
                   [np.full(d, 0.5)] * M, [o.noise for o in st.objectives], [o.y_mean for o in st.objectives], [o.y_std for o in st.objectives],
                   np.zeros(d), np.ones(d))
    X = rng.random((P, d))
    Xd = L.DeviceArray((P, d)).upload(X)
    md, vd = L.DeviceArray((P, M)), L.DeviceArray((P, M))
    lib, ctx = L.load_library(), L.context()
    prior = np.array([(o.constant + o.noise) * o.y_std**2 for o in st.objectives])
    ref = None
    for name, prec in (("fp64", L.GP_FP64), ("tensor", L.GP_TENSOR)):
        L.profile_enable(True)
        ms = timed(lambda: L._check(lib.dmo_gp_predict(ctx, h._h, Xd.ptr, P, md.ptr, vd.ptr, prec), "gp"), reps=2)
        rep = L.profile_report()
        L.profile_enable(False)
        mean, var = md.download(), vd.download()
        if ref is None:
            ref = (mean, var)
            err = ""
        else:
            err = f" | vs fp64: var err/prior {np.max(np.abs(var - ref[1]) / prior):.2e}, mean err {np.max(np.abs(mean - ref[0])):.2e}"
        parts = ", ".join(f"{k} {v[0] / v[1]:.3f}" for k, v in rep.items())
        print(f"gp {name} (DMO_GP_TC={os.environ.get('DMO_GP_TC', 'default')}): total {ms:.3f} ms [{parts}]{err}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gp":
        gp_sweep()
        sys.exit(0)
    main()
