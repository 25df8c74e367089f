"""Oracle: non-dominated ranking (rows A1/A2 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/dda.py``:
  * ``dominance_degree_matrix``  -> dda.py:13-47   (D[i,j] = #objectives with y_i <= y_j)
  * ``dda_ens`` / ``dda_insert`` -> dda.py:97-152  (ENS-SS insertion in argsort(Y[:,0]) order)
  * ``dda_non_dominated_sort``   -> dda.py:50-94   (front peeling on D)

and adds ``rank_canonical`` -- the canonical Pareto front index (length of the
longest domination chain ending at a point).  The CUDA kernel implements the
canonical rank; it equals ``dda_ens`` whenever objective 0 is tie-free
(SURVEY.md section 8a row A2), which tests assert.
"""

import numpy as np


def dominance_degree_matrix(Y):
    """D[i, j] = number of objectives k with Y[i,k] <= Y[j,k]  (dda.py:13-47).

    The reference builds this from per-objective comparison matrices filled in
    sorted order; the closed form is the same integer matrix.
    """
    Y = np.asarray(Y)
    return (Y[:, None, :] <= Y[None, :, :]).sum(axis=2).astype(np.intp)


def _zero_identical(D, d):
    """dda.py:108-115: identical vectors are made mutually non-dominating."""
    same = (D == d) & (D.T == d)
    D = D.copy()
    D[same] = 0
    return D


def dda_ens(Y):
    """Faithful restatement of dda.py:97-152 (the rank every sortMO uses).

    Points are inserted in ``np.argsort(Y[:, 0])`` order into the first front
    none of whose *current* members dominates them.
    """
    Y = np.asarray(Y)
    n, d = Y.shape
    D = _zero_identical(dominance_degree_matrix(Y), d)
    fronts = []
    rank = np.zeros(n, dtype=np.intp)
    for s in np.argsort(Y[:, 0]):
        placed = False
        for k, front in enumerate(fronts):
            if not np.any(D[front, s] == d):
                front.append(s)
                rank[s] = k
                placed = True
                break
        if not placed:
            fronts.append([s])
            rank[s] = len(fronts) - 1
    return rank


def dominates_matrix(Y):
    """dom[i, j] = True iff i dominates j (all <=, not identical)."""
    Y = np.asarray(Y)
    le = (Y[:, None, :] <= Y[None, :, :]).all(axis=2)
    eq = (Y[:, None, :] == Y[None, :, :]).all(axis=2)
    return le & ~eq


def rank_canonical(Y, block=2048):
    """Canonical non-dominated rank by front peeling; O(F * n^2 / block) memory-light.

    Equals dda.py:50-94 (``dda_non_dominated_sort``) for every input and
    ``dda_ens`` whenever objective 0 has no ties.
    """
    Y = np.asarray(Y, dtype=np.float64)
    n = Y.shape[0]
    rank = np.full(n, -1, dtype=np.int64)
    alive = np.arange(n)
    k = 0
    while alive.size:
        Ya = Y[alive]
        dominated = np.zeros(alive.size, dtype=bool)
        for s in range(0, alive.size, block):
            Yb = Ya[s : s + block]
            le = (Ya[:, None, :] <= Yb[None, :, :]).all(axis=2)
            eq = (Ya[:, None, :] == Yb[None, :, :]).all(axis=2)
            dominated[s : s + block] = (le & ~eq).any(axis=0)
        rank[alive[~dominated]] = k
        alive = alive[dominated]
        k += 1
    return rank


def rank_chain_dp(Y):
    """Canonical rank as longest-chain DP over the lexicographic order.

    This is the formulation the CUDA kernel uses (rank_i = 1 + max rank of the
    dominators of i, processed in lexicographic order); kept here so the DP
    itself is pinned against peeling on the CPU.
    """
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    order = np.lexsort(tuple(Y[:, k] for k in range(d - 1, -1, -1)))
    Ys = Y[order]
    r = np.zeros(n, dtype=np.int64)
    for i in range(1, n):
        le = (Ys[:i] <= Ys[i]).all(axis=1)
        eq = (Ys[:i] == Ys[i]).all(axis=1)
        dom = le & ~eq
        if dom.any():
            r[i] = r[:i][dom].max() + 1
    out = np.empty(n, dtype=np.int64)
    out[order] = r
    return out
