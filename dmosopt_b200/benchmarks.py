"""Vectorised benchmark objective functions on the GPU.

Same names and formulas as ``dmosopt/benchmarks/moo_benchmarks.py`` (dtlz1 :21, dtlz2 :59, dtlz3 :97, dtlz4 :136,
dtlz5 :174, dtlz7 :218, wfg4 :335) plus the two example objectives ZDT1 / ZDT3 (examples/example_dmosopt_zdt1.py:9-20,
examples/example_dmosopt_zdt3.py:9-21).  The reference functions take ONE decision vector; these take one (returning a
1-D array, drop-in) or a whole (n, n_var) matrix (returning (n, n_obj)), evaluated by dmo_benchmark_eval.
"""

import numpy as np

from . import _lib


def _eval(name, x, n_obj, **kw):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return _lib.benchmark_eval(name, x[None, :], n_obj, **kw)[0]
    return _lib.benchmark_eval(name, x, n_obj, **kw)


def zdt1(x):
    return _eval("zdt1", x, 2)


def zdt3(x):
    return _eval("zdt3", x, 2)


def dtlz1(x, n_obj=3):
    return _eval("dtlz1", x, n_obj)


def dtlz2(x, n_obj=3):
    return _eval("dtlz2", x, n_obj)


def dtlz3(x, n_obj=3):
    return _eval("dtlz3", x, n_obj)


def dtlz4(x, n_obj=3, alpha=100.0):
    return _eval("dtlz4", x, n_obj, alpha=alpha)


def dtlz5(x, n_obj=3):
    return _eval("dtlz5", x, n_obj)


def dtlz7(x, n_obj=3):
    return _eval("dtlz7", x, n_obj)


def wfg4(x, n_obj=3, k=None):
    if k is not None and k != n_obj - 1:
        raise NotImplementedError("wfg4: only the reference's default position parameter k = n_obj - 1 is built")
    return _eval("wfg4", x, n_obj)
