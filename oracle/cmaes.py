"""Oracle: MO-CMA-ES sampling and selection (rows A13-A15 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/CMAES.py``:
  * local ``sortMO``      -> CMAES.py:455-486 (perm = lexsort(rank), rank unsorted)
  * ``_select``           -> CMAES.py:167-229 (whole fronts, then split the overflowing front with
                             HypervolumeImprovement(ref = max(y) + 1, nds=True).do(chosen_y, mid_y, ones, k))
  * ``generate_strategy`` -> CMAES.py:231-271 (x = parent + sigma * (A @ z), then the global rescale)
  * ``updateCholesky``    -> CMAES.py:489-537
"""

import numpy as np

from .dda import dda_ens
from .hv import select_candidates


def sort_mo(y, rank_fn=dda_ens):
    """CMAES.py:455-486 without x metrics: (perm, rank) with rank in original order."""
    rank = rank_fn(np.asarray(y))
    return np.lexsort([rank]), rank


def select(candidates_y, popsize, rank_fn=dda_ens):
    """CMAES.py:167-229.  Returns (chosen, not_chosen, rank) boolean masks over the candidates.

    NB the reference indexes fronts through ``order_inv[argwhere(rank == r)]`` (:190) where
    ``order_inv = argsort(order)``; i.e. front r is mapped through the inverse permutation.
    That is reproduced literally here.
    """
    y = np.asarray(candidates_y)
    n = y.shape[0]
    if n <= popsize:
        return np.ones(n, dtype=bool), np.zeros(n, dtype=bool), None
    order, rank = sort_mo(y, rank_fn)
    order_inv = np.argsort(order)
    chosen = np.zeros(n, dtype=bool)
    not_chosen = np.zeros(n, dtype=bool)
    mid = None
    full = False
    count = 0
    for r in range(int(rank.max()) + 1):
        fr = order_inv[np.flatnonzero(rank == r)]
        if count + len(fr) <= popsize and not full:
            chosen[fr] = True
            count += len(fr)
        elif mid is None and count < popsize:
            mid = fr.copy()
            full = True
        else:
            not_chosen[fr] = True
    k = popsize - count
    if k > 0:
        ref = np.max(y, axis=0) + 1
        if count > 0:
            F = y[chosen]
            r0 = rank_fn(F)
            front = F[r0 == 0] if np.any(r0 == 0) else F
            sel, _ = select_candidates(front, y[mid], np.ones_like(y[mid]), ref, k)
        else:
            sel = np.arange(k)
        chosen[mid[sel]] = True
        m = np.ones(len(mid), dtype=bool)
        m[sel] = False
        not_chosen[mid[m]] = True
    return chosen, not_chosen, rank


def sample(parents_x, sigmas, A, p_idx, arz, bounds):
    """CMAES.py:263-271: individuals = x_p + sigma_p * (A_p @ z), then global rescale into bounds."""
    ind = parents_x[p_idx] + sigmas[p_idx] * np.einsum("ijk,ik->ij", A[p_idx], arz)
    xrng = bounds[:, 1] - bounds[:, 0]
    return (ind / np.max(np.abs(ind))) * xrng + bounds[:, 0]


def update_cholesky(A, Ainv, z, psucc, pc, cc, ccov, pthresh):
    """CMAES.py:489-537 rank-one update of the Cholesky factor and its inverse."""
    if psucc < pthresh:
        pc = (1.0 - cc) * pc + np.sqrt(cc * (2.0 - cc)) * z
        alpha = 1.0 - ccov
    else:
        pc = (1.0 - cc) * pc
        alpha = (1.0 - ccov) + ccov * cc * (2.0 - cc)
    beta = ccov
    w = Ainv @ pc
    if w.max() > 1e-20:
        wA = w @ Ainv
        a = np.sqrt(alpha)
        n2 = np.sum(w**2)
        root = np.sqrt(1 + beta / alpha * n2)
        b = a / n2 * (root - 1)
        A = a * A + b * np.outer(pc, w)
        c = 1.0 / (a * n2) * (1.0 - 1.0 / root)
        Ainv = (1.0 / a) * Ainv - c * np.outer(w, wA)
    return A, Ainv, pc
