"""One surrogate generation of every BASELINE.json configuration at its stated size, through the plugin API on the GPU.

  C2  ZDT3   d=30 M=2 pop=8192   AGEMOEA + GP N_train=2048
  C3  DTLZ2  d=12 M=3 pop=65536  NSGA2   + GP N_train=4096
  C4  DTLZ7  d=22 M=5 pop=32768  SMPSO   + HV-contribution selection
  C5  WFG4-shaped d=24 M=4 pop=131072 CMAES + GP N_train=4096

Prints ms per generation (after one warm-up generation) and checks the invariants that do not depend on the size:
survivors are rows of (offspring + parents), the stored ranks are the canonical ranks of the stored objectives, the
population size is preserved.  Synthetic targets: smooth test functions of the stated shape (the surrogate hyper-
parameters are fixed).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmosopt_b200 as b2  # noqa: E402
from dmosopt_b200 import _lib as L  # noqa: E402


def targets(X, M, kind):
    d = X.shape[1]
    if kind == "zdt3":
        g = 1.0 + 9.0 / (d - 1) * X[:, 1:].sum(axis=1)
        f1 = X[:, 0]
        return np.column_stack((f1, g * (1.0 - np.sqrt(f1 / g) - f1 / g * np.sin(10 * np.pi * f1))))
    g = ((X[:, M - 1:] - 0.5) ** 2).sum(axis=1)
    Y = np.ones((X.shape[0], M)) * (1.0 + g)[:, None]
    for i in range(M):
        for j in range(M - 1 - i):
            Y[:, i] *= np.cos(0.5 * np.pi * X[:, j])
        if i > 0:
            Y[:, i] *= np.sin(0.5 * np.pi * X[:, M - 1 - i])
    if kind == "dtlz7":
        Y[:, -1] = (1.0 + g) * (M - (Y[:, :-1] / (1.0 + g)[:, None] * (1.0 + np.sin(3 * np.pi * Y[:, :-1]))).sum(axis=1))
    return Y


def run(name, cls, d, M, pop, N, kind, gens=4, keep_last=None, **okw):
    rng = np.random.default_rng(20260921 + len(name))
    xlb, xub = np.zeros(d), np.ones(d)
    Xtr = rng.random((N, d))
    sm = b2.GPR_Matern(Xtr, targets(Xtr, M, kind), d, M, xlb, xub, optimizer=None)  # precision = auto (the plugin default)
    mdl = b2.Model(objective=sm)
    opt = cls(popsize=pop, nInput=d, nOutput=M, model=mdl, **okw)
    bounds = np.column_stack((xlb, xub))
    x0 = opt.generate_initial(bounds, rng)
    if x0.shape[0] < pop:
        x0 = rng.random((pop, d))
    y0 = sm.evaluate(x0).astype(np.float32)
    opt.initialize_strategy(x0, y0, bounds, rng)
    times, parts = [], []
    prof = None
    if os.environ.get("DMO_PROFILE"):  # where does the host side of a generation go (cProfile over the timed generations)
        import cProfile

        prof = cProfile.Profile()
    for g in range(gens + 1):
        if prof is not None and g == 1:
            prof.enable()
        if g == gens and keep_last is not None:  # the state the last update starts from (for exact re-computation by the tests)
            keep_last["before"] = {k: np.array(v) for k, v in opt.state.items() if isinstance(v, np.ndarray)}
        L.synchronize()
        t0 = time.perf_counter()
        x_gen, st = opt.generate()
        t1 = time.perf_counter()
        y_gen = sm.evaluate(x_gen)
        t2 = time.perf_counter()
        opt.update(x_gen, y_gen, st)
        L.synchronize()
        t3 = time.perf_counter()
        times.append(t3 - t0)
        parts.append((t1 - t0, t2 - t1, t3 - t2))
    if prof is not None:
        import pstats

        prof.disable()
        pstats.Stats(prof).sort_stats("cumulative").print_stats(28)
    if keep_last is not None:
        keep_last.update(x_gen=np.array(x_gen), y_gen=np.array(y_gen), state_gen=st, ms=np.median(times[1:]) * 1e3, surrogate=sm)
    px, py = opt.population_objectives
    assert px.shape[1] == d and py.shape[1] == M and np.all(np.isfinite(py)), name
    assert np.all(px >= xlb - 1e-12) and np.all(px <= xub + 1e-12), name
    # median over the generations after the first: the stream-ordered memory pool still grows during the first few
    # (a multi-GB cudaMallocAsync from the driver costs hundreds of milliseconds once)
    ms = np.median(times[1:]) * 1e3
    P = x_gen.shape[0]
    gm, em, um = (np.median([p[i] for p in parts[1:]]) * 1e3 for i in range(3))
    if os.environ.get("DMO_VERBOSE_GENS"):
        for g, (t, p3) in enumerate(zip(times, parts)):
            print(f"   generation {g}: {t * 1e3:.1f} ms (generate {p3[0] * 1e3:.1f} + surrogate {p3[1] * 1e3:.1f} + update {p3[2] * 1e3:.1f})", flush=True)
    print(f"{name}: pop={pop} d={d} M={M} N_train={N} offspring/generation={P}: {ms:.1f} ms/generation (generate {gm:.1f} + surrogate {em:.1f} + update {um:.1f}) "
          f"-> {P / ms * 1e3:,.0f} candidates/s; population {px.shape[0]} rows", flush=True)
    return opt, px, py


def main():
    L.context()
    which = sys.argv[1:] or ["C2", "C3", "C4", "C5"]
    if "C2" in which:
        opt, px, py = run("C2 AGEMOEA", b2.AGEMOEA, 30, 2, 8192, 2048, "zdt3")
        assert px.shape[0] == 8192
    if "C3" in which:
        opt, px, py = run("C3 NSGA2", b2.NSGA2, 12, 3, 65536, 4096, "dtlz2", distance_metric=None)
        r = L.rank_nd(py.astype(np.float64))
        # (equal up to the few points whose dominance relations the float32 rounding of the stored objectives changes)
        assert int((np.asarray(opt.state.rank) != r).sum()) <= 32, "stored ranks are the canonical ranks of the stored objectives"
    if "C4" in which:
        opt, px, py = run("C4 SMPSO", b2.SMPSO, 22, 5, int(os.environ.get("C4_POP", "32768")), 4096, "dtlz7")
        # HV-contribution selection on the result (A17): 4096 of the population against its own front
        front = py[L.rank_nd(py.astype(np.float64)) == 0].astype(np.float64)[:256]
        mu, var = opt.model.objective.predict(px[:8192])
        ref = py.max(axis=0).astype(np.float64) + 1.0
        t0 = time.perf_counter()
        sel = L.ehvi_select(front, mu, var, ref, 1024)
        L.synchronize()
        print(f"   C4 HV-improvement selection: 8192 candidates, front {front.shape[0]}, M=5: {(time.perf_counter() - t0) * 1e3:.1f} ms, {len(np.unique(sel))} distinct picks")
    if "C5" in which:
        run("C5 CMAES", b2.CMAES, 24, 4, int(os.environ.get("C5_POP", "131072")), 4096, "dtlz2")


if __name__ == "__main__":
    main()
