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


if __name__ == "__main__":
    main()
