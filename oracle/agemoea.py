"""Oracle: AGE-MOEA survival (row A11 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/AGEMOEA.py``:
  * local ``sortMO``            -> AGEMOEA.py:261-272
  * ``normalize``               -> AGEMOEA.py:275-315
  * ``minkowski_distances``     -> AGEMOEA.py:318-321
  * ``get_geometry``            -> AGEMOEA.py:324-339
  * ``point_2_line_distance``   -> AGEMOEA.py:342-351
  * ``find_corner_solutions``   -> AGEMOEA.py:354-374
  * ``survival_score``          -> AGEMOEA.py:377-430
  * ``environmental_selection`` -> AGEMOEA.py:433-501

The reference orders rows with an unstable ``rank.argsort()`` (:267), so the row
order inside one front is not reproducible; a stable order is used here and
parity is asserted on the selected *set* (rows sorted lexicographically).
"""

import numpy as np

from .dda import dda_ens


def line_distance(P, direction):
    """Distance of each row of P to the line through the origin along ``direction`` (:342-351)."""
    b = np.asarray(direction, dtype=np.float64)
    t = (P @ b) / (b @ b)
    return np.linalg.norm(P - t[:, None] * b[None, :], axis=1)


def corner_solutions(front):
    """AGEMOEA.py:354-374: for each axis (1e-6 + e_i), the unselected point closest to it."""
    m, n = front.shape
    if m <= n:
        return np.arange(m)
    W = 1e-6 + np.eye(n)
    taken = np.zeros(m, dtype=bool)
    idx = np.zeros(n, dtype=int)
    for i in range(n):
        dd = line_distance(front, W[i])
        dd[taken] = np.inf
        idx[i] = int(np.argmin(dd))
        taken[idx[i]] = True
    return idx


def hyperplane_normalization(front, extreme):
    """AGEMOEA.py:275-315: axis intercepts of the hyperplane through the extreme points."""
    n = front.shape[1]
    if len(extreme) != len(np.unique(extreme)):
        return np.max(front, axis=0)
    try:
        h = np.linalg.solve(front[extreme], np.ones(n))
    except Exception:
        h = np.array([np.nan])
    if np.any(np.isnan(h)) or np.any(np.isinf(h)) or np.any(h < 0):
        nz = np.max(front, axis=0)
    else:
        with np.errstate(divide="ignore"):
            nz = 1.0 / h
        if np.any(np.isnan(nz)) or np.any(np.isinf(nz)):
            nz = np.max(front, axis=0)
    nz = np.array(nz, dtype=np.float64)
    nz[np.isclose(nz, 0.0, rtol=1e-4, atol=1e-4)] = 1.0
    return nz


def geometry_p(front, extreme):
    """AGEMOEA.py:324-339: L_p curvature from the most central point."""
    n = front.shape[1]
    dd = line_distance(front, np.ones(n))
    dd[extreme] = np.inf
    i = int(np.argmin(dd))
    with np.errstate(divide="ignore", invalid="ignore"):
        p = np.log(n) / np.log(1.0 / np.mean(front[i]))
    if np.isnan(p) or p <= 0.1:
        return 1.0
    if p > 20:
        return 20.0
    return float(p)


def minkowski(A, B, p):
    """AGEMOEA.py:318-321: out[j, i] = ||A[i] - B[j]||_p."""
    return np.power(np.power(np.abs(A[None, :, :] - B[:, None, :]), p).sum(axis=2), 1.0 / p)


def survival_score(yfront_raw, ideal_point):
    """AGEMOEA.py:377-430 for one front given as its own array.  Returns (normalization, p, crowd)."""
    m, n = yfront_raw.shape
    crowd = np.zeros(m)
    if m < n:
        nz = np.max(yfront_raw, axis=0).astype(np.float64)
        nz[np.isclose(nz, 0.0, rtol=1e-4, atol=1e-4)] = 1.0
        return nz, 1.0, crowd
    yf = yfront_raw - ideal_point
    extreme = corner_solutions(yf)
    nz = hyperplane_normalization(yf, extreme)
    yn = yf / nz
    p = geometry_p(yn, extreme)
    crowd[extreme] = np.inf
    selected = np.zeros(m, dtype=bool)
    selected[extreme] = True
    nn = np.linalg.norm(yn, p, axis=1)
    dist = minkowski(yn, yn, p) / nn[:, None]  # dist[s, r] = d(s, r) / nn[s]
    remaining = [i for i in range(m) if not selected[i]]
    # two smallest normalised distances from each remaining point to the selected set
    while remaining:
        sel = np.flatnonzero(selected)
        D = dist[np.ix_(sel, remaining)].T  # (remaining, selected)
        if D.shape[1] > 1:
            part = np.partition(D, 1, axis=1)[:, :2]
            score = part.sum(axis=1)
        else:
            score = D[:, 0]
        j = int(np.argmax(score))
        best = remaining.pop(j)
        selected[best] = True
        crowd[best] = score[j]
    return nz, p, crowd


def environmental_selection(x, y, pop, rank_fn=dda_ens):
    """AGEMOEA.py:433-501.  Returns (xs, ys, rank, crowd_dist float32) of the survivors."""
    x = np.asarray(x)
    y = np.asarray(y)
    rank = rank_fn(y)
    o = rank.argsort(kind="stable")
    xs, ys, rank = x[o], y[o], rank[o]
    rmax = int(rank.max())
    n = ys.shape[0]
    crowd = np.zeros(n, dtype=np.float32)
    selected = np.zeros(n, dtype=bool)
    f1 = np.flatnonzero(rank == 0)
    ideal = np.min(ys[f1], axis=0)
    nz, p, c1 = survival_score(ys[f1], ideal)
    crowd[f1] = c1
    count = len(f1)
    if count < pop:
        selected[f1] = True
        for r in range(1, rmax + 1):
            fr = np.flatnonzero(rank == r)
            yn = ys[fr] / nz
            # NB the reference measures normalised points against the *un-normalised* ideal point (:466-471)
            crowd[fr] = 1.0 / minkowski(yn, ideal[None, :], p)[0]
            if count + len(fr) < pop:
                selected[fr] = True
                count += len(fr)
            else:
                perm = np.lexsort([-crowd[fr]])
                selected[fr[perm[: pop - count]]] = True
                break
    else:
        perm = np.lexsort([-crowd[f1]])
        selected[f1[perm[:pop]]] = True
    return xs[selected].copy(), ys[selected].copy(), rank[selected].copy(), crowd[selected].copy()
