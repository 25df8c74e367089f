"""Host-side profile of one plugin-API generation at the BASELINE shape (where does e2e time go beyond the kernels?).

Prints per-call wall times of the plugin step's four calls, the library's device timers, and a cProfile listing.
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dmosopt_b200 as b2  # noqa: E402
from dmosopt_b200 import _lib as L  # noqa: E402
from dmosopt_b200.indicators import Hypervolume  # noqa: E402


def main():
    pop, d, M, N = 65536, 30, 3, 4096
    L.context()
    w = bench.workload(pop, d, M, N)
    sm = b2.GPR_Matern(w["Xtr"], w["Ytr"], d, M, w["xlb"], w["xub"], optimizer=None)  # precision = auto (the plugin default)
    mdl = b2.Model(objective=sm)
    y0 = sm.evaluate(w["X0"]).astype(np.float32)
    ref = y0.max(axis=0).astype(np.float64) + 0.1 * (y0.max(axis=0) - y0.min(axis=0))
    opt = b2.NSGA2(popsize=pop, nInput=d, nOutput=M, model=mdl, distance_metric=None)
    opt.initialize_strategy(w["X0"], y0, np.column_stack((w["xlb"], w["xub"])), np.random.default_rng(0))
    hv = Hypervolume(ref_point=ref)
    acc = {}

    def tick(name, t0):
        L.synchronize()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0

    def step():
        t = time.perf_counter()
        x_gen, st = opt.generate()
        tick("generate", t)
        t = time.perf_counter()
        y_gen, y_var = sm.predict(x_gen)
        tick("predict", t)
        t = time.perf_counter()
        opt.update(x_gen, y_gen, st)
        tick("update", t)
        t = time.perf_counter()
        _, py = opt.population_objectives
        tick("population_objectives", t)
        t = time.perf_counter()
        hv.do(py.astype(np.float64))
        tick("hypervolume", t)

    for _ in range(3):
        step()
    acc.clear()
    K = 5
    h0, d0 = L.transfer_bytes()
    L.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    L.synchronize()
    tot = time.perf_counter() - t0
    rep = L.profile_report()
    L.profile_enable(False)
    h1, d1 = L.transfer_bytes()
    print(f"plugin step: {tot / K * 1e3:.2f} ms  (h2d {(h1 - h0) / K / 1e6:.1f} MB, d2h {(d1 - d0) / K / 1e6:.1f} MB per step)")
    for k, v in acc.items():
        print(f"  {k:24s} {v / K * 1e3:8.3f} ms")
    print("device timers (ms per step):", {k: round(v[0] / K, 3) for k, v in rep.items()})
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
