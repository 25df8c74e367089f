"""Oracle: benchmark objective functions, vectorised over rows (test infrastructure only, see oracle/__init__.py).

Restates dmosopt/benchmarks/moo_benchmarks.py (dtlz1 :21-56, dtlz2 :59-94, dtlz3 :97-133, dtlz4 :136-171, dtlz5 :174-215,
dtlz7 :218-253, wfg4 :335-375) and the example objectives examples/example_dmosopt_zdt1.py:9-20 /
examples/example_dmosopt_zdt3.py:9-21; pinned to the reference by tests/golden/trs.npz (rows evaluated one at a time by
the reference itself).
"""

import numpy as np


def _product_form(v, M, scale, c, s):
    f = np.empty((v.shape[0], M))
    for i in range(M):
        fi = np.array(scale, dtype=np.float64, copy=True) * np.ones(v.shape[0])
        for j in range(M - i - 1):
            fi = fi * c(v[:, j])
        if i > 0:
            fi = fi * s(v[:, M - i - 1])
        f[:, i] = fi
    return f


def _g_sphere(X, M):
    k = X.shape[1] - M + 1
    return np.sum((X[:, -k:] - 0.5) ** 2, axis=1)


def zdt1(X):
    X = np.atleast_2d(X)
    g = 1.0 + 9.0 / (X.shape[1] - 1) * np.sum(X[:, 1:], axis=1)
    return np.column_stack((X[:, 0], g * (1.0 - np.sqrt(X[:, 0] / g))))


def zdt3(X):
    X = np.atleast_2d(X)
    g = 1.0 + 9.0 / (X.shape[1] - 1) * np.sum(X[:, 1:], axis=1)
    f1 = X[:, 0]
    # examples/example_dmosopt_zdt3.py:18-20: f2 = g h - j with j = (x0 / g) sin(10 pi x0) (j is NOT scaled by g there)
    return np.column_stack((f1, g * (1.0 - np.sqrt(f1 / g)) - (f1 / g) * np.sin(10 * np.pi * f1)))


def dtlz1(X, M=3):
    X = np.atleast_2d(X)
    k = X.shape[1] - M + 1
    xm = X[:, -k:] - 0.5
    g = 100 * (k + np.sum(xm**2 - np.cos(20 * np.pi * xm), axis=1))
    return _product_form(X, M, 0.5 * (1 + g), lambda v: v, lambda v: 1 - v)


def dtlz2(X, M=3):
    X = np.atleast_2d(X)
    return _product_form(X, M, 1 + _g_sphere(X, M), lambda v: np.cos(v * np.pi / 2), lambda v: np.sin(v * np.pi / 2))


def dtlz3(X, M=3):
    X = np.atleast_2d(X)
    k = X.shape[1] - M + 1
    xm = X[:, -k:] - 0.5
    g = 100 * (k + np.sum(xm**2 - np.cos(20 * np.pi * xm), axis=1))
    return _product_form(X, M, 1 + g, lambda v: np.cos(v * np.pi / 2), lambda v: np.sin(v * np.pi / 2))


def dtlz4(X, M=3, alpha=100.0):
    X = np.atleast_2d(X)
    return _product_form(X, M, 1 + _g_sphere(X, M), lambda v: np.cos(v**alpha * np.pi / 2), lambda v: np.sin(v**alpha * np.pi / 2))


def dtlz5(X, M=3):
    X = np.atleast_2d(X)
    g = _g_sphere(X, M)
    th = np.empty((X.shape[0], M - 1))
    th[:, 0] = X[:, 0] * np.pi / 2
    for i in range(1, M - 1):
        th[:, i] = (1 + 2 * g * X[:, i]) / (2 * (1 + g)) * np.pi / 2
    return _product_form(th, M, 1 + g, np.cos, np.sin)


def dtlz7(X, M=3):
    X = np.atleast_2d(X)
    k = X.shape[1] - M + 1
    g = 1 + 9 * np.mean(X[:, -k:], axis=1)
    f = np.empty((X.shape[0], M))
    f[:, :-1] = X[:, : M - 1]
    h = M - np.sum(f[:, :-1] / (1 + g)[:, None] * (1 + np.sin(3 * np.pi * f[:, :-1])), axis=1)
    f[:, -1] = (1 + g) * h
    return f


def wfg4(X, M=3):
    X = np.atleast_2d(X)
    d = X.shape[1]
    ll = d - (M - 1)
    y = X / (2 * np.arange(1, d + 1))
    t1 = y + 0.35 - 0.15 * np.cos(10 * np.pi * y - 5)
    xv = np.empty((X.shape[0], M))
    with np.errstate(invalid="ignore"):
        for i in range(M - 1):
            w = t1[:, i * ll : (i + 1) * ll]
            xv[:, i] = w.mean(axis=1) if w.shape[1] else np.nan
        xv[:, -1] = t1[:, -ll:].mean(axis=1)
    f = _product_form(xv, M, 1.0, lambda v: 1 - np.cos(v * np.pi / 2), lambda v: 1 - np.sin(v * np.pi / 2))
    return f * (1 + np.arange(1, M + 1))
