"""Oracle: distance metrics (rows A3/A4 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/indicators.py``:
  * ``crowding_distance_metric``  -> indicators.py:12-51
  * ``euclidean_distance_metric`` -> indicators.py:54-62
"""

import numpy as np


def crowding_distance_loops(Y):
    """Literal restatement of indicators.py:12-51 (Python loops, small n only)."""
    Y = np.asarray(Y)
    n, d = Y.shape
    lb = np.min(Y, axis=0, keepdims=True)
    ub = np.max(Y, axis=0, keepdims=True)
    if n == 1:
        return np.array([1.0])
    rng = ub - lb
    rng[rng == 0.0] = 1.0
    U = (Y - lb) / rng
    D = np.zeros(n)
    DS = np.zeros((n, d))
    idx = U.argsort(axis=0, kind="stable")
    US = np.zeros((n, d))
    for j in range(d):
        US[:, j] = U[idx[:, j], j]
    DS[0, :] = 1.0
    DS[n - 1, :] = 1.0
    for i in range(1, n - 1):
        for j in range(d):
            DS[i, j] = US[i + 1, j] - US[i - 1, j]
    for i in range(n):
        for j in range(d):
            D[idx[i, j]] += DS[i, j]
    D[np.isnan(D)] = 0.0
    return D


def crowding_distance_metric(Y):
    """Vectorised, bit-identical restatement of indicators.py:12-51.

    Global min-max normalisation (zero range -> 1.0, :26-29); per-objective
    argsort; interior contribution = next - prev, both ends = 1.0 (:39-40);
    the reference accumulates D[idx[i, j]] in (i, j) loop order (:46-48), so a
    point's M contributions are added in order of (sorted position, objective)
    -- reproduced here so the float64 result is identical to the last bit.
    The reference's ``argsort`` is numpy's default (unstable) sort; a stable
    order is used here, identical whenever an objective column is tie-free.
    """
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    if n == 1:
        return np.array([1.0])
    lb = np.min(Y, axis=0, keepdims=True)
    ub = np.max(Y, axis=0, keepdims=True)
    rng = ub - lb
    rng[rng == 0.0] = 1.0
    U = (Y - lb) / rng
    idx = U.argsort(axis=0, kind="stable")
    pos = np.empty((n, d), dtype=np.int64)
    contrib = np.empty((n, d))
    ar = np.arange(n)
    for j in range(d):
        us = U[idx[:, j], j]
        ds = np.empty(n)
        ds[0] = 1.0
        ds[n - 1] = 1.0
        ds[1 : n - 1] = us[2:] - us[: n - 2]
        pos[idx[:, j], j] = ar
        contrib[idx[:, j], j] = ds
    # add each point's contributions in (position, objective) order
    key = pos * d + np.arange(d)[None, :]
    o = np.argsort(key, axis=1, kind="stable")
    c = np.take_along_axis(contrib, o, axis=1)
    D = np.zeros(n)
    for j in range(d):
        D = D + c[:, j]
    D[np.isnan(D)] = 0.0
    return D


def euclidean_distance_metric(Y):
    """indicators.py:54-62: row norm of the min-max-normalised objectives."""
    Y = np.asarray(Y, dtype=np.float64)
    lb = np.min(Y, axis=0)
    ub = np.max(Y, axis=0)
    rng = ub - lb
    rng[rng == 0.0] = 1.0
    U = (Y - lb) / rng
    return np.sqrt(np.sum(U**2, axis=1))
