"""Oracle: exact hypervolume and EHVI candidate selection (rows A16/A17 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Reference:
  * ``dmosopt/hv.py:123-189``  AdaptiveHyperVolume.compute_hypervolume(..., 'box'):
    keeps only points strictly inside ``ref`` (:159-162) then calls the box algorithm.
  * ``dmosopt/hv_box_decomposition.py:86-304``  HyperVolumeBoxDecomposition.compute_hypervolume
    (Lacour-Klamroth-Fonseca local-upper-bound decomposition).
  * ``dmosopt/hv_box_decomposition.py:306-437``  select_candidates / _compute_batch_ehvi /
    _decompose_dominated_space (the "HV contribution" selection used by CMAES / TRS through
    ``indicators.HypervolumeImprovement._do``, indicators.py:295-313).

``hypervolume`` below is the *true* hypervolume (minimisation), computed by
dimension sweep -- an independent algorithm, not a transliteration of the box
decomposition.  It equals the reference for every input whose first M-1
objectives are strictly positive; the reference silently loses volume
otherwise because its dummy defining points sit at 0 (hv_box_decomposition.py
:136-145, :228) -- golden fixtures therefore use
positive objectives, and the divergence case is pinned as a documented
difference in tests/test_oracle_golden.py.
"""

import numpy as np
from scipy.stats import norm


def inside_ref(points, ref):
    """hv.py:159-162: keep points with ref > p in every objective."""
    points = np.asarray(points, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return points[np.all(ref > points, axis=1)]


def _nondominated(P):
    """Drop weakly dominated points and exact duplicates (does not change the HV)."""
    n = P.shape[0]
    if n <= 1:
        return P
    P = np.unique(P, axis=0)
    le = (P[:, None, :] <= P[None, :, :]).all(axis=2)
    np.fill_diagonal(le, False)
    return P[~le.any(axis=0)]


def _hv2d(P, ref):
    """Staircase sweep: sort by f0, accumulate strips under the running min of f1."""
    o = np.argsort(P[:, 0], kind="stable")
    x = P[o, 0]
    y = P[o, 1]
    ymin = np.minimum.accumulate(y)
    xn = np.append(x[1:], ref[0])
    return float(np.sum((xn - x) * (ref[1] - ymin)))


def _hv_rec(P, ref):
    d = P.shape[1]
    if P.shape[0] == 0:
        return 0.0
    if d == 1:
        return float(ref[0] - P[:, 0].min())
    if d == 2:
        return _hv2d(P, ref)
    # slice along the last objective
    o = np.argsort(P[:, -1], kind="stable")
    P = P[o]
    z = P[:, -1]
    zn = np.append(z[1:], ref[-1])
    total = 0.0
    for i in range(P.shape[0]):
        h = zn[i] - z[i]
        if h > 0.0:
            total += h * _hv_rec(_nondominated(P[: i + 1, :-1]), ref[:-1])
    return total


def hypervolume(points, ref):
    """True hypervolume dominated by ``points`` and bounded by ``ref`` (minimisation)."""
    ref = np.asarray(ref, dtype=np.float64)
    P = inside_ref(points, ref)
    if P.shape[0] == 0:
        return 0.0
    return _hv_rec(_nondominated(P), ref)


# ---------------------------------------------------------------------------
# EHVI candidate selection (hv_box_decomposition.py:306-437)
# ---------------------------------------------------------------------------


def decompose_boxes(front, ref):
    """hv_box_decomposition.py:418-437: boxes between consecutive f0-sorted front points.

    lower = (-inf, front sorted by f0), upper = (front sorted by f0, ref); only boxes
    with upper > lower in *every* objective are kept.  (For a mutually
    non-dominated front this leaves the two end boxes -- the reference's
    heuristic box set, reproduced as is.)
    """
    front = np.asarray(front, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    n, d = front.shape
    o = np.argsort(front[:, 0])
    sf = front[o]
    lower = np.full((n + 1, d), -np.inf)
    upper = np.full((n + 1, d), np.inf)
    lower[1:] = sf
    upper[:-1] = sf
    upper[-1] = ref
    valid = np.all(upper > lower, axis=1)
    return lower[valid], upper[valid]


def batch_ehvi(lower, upper, means, variances):
    """hv_box_decomposition.py:353-416: score_c = sum_b prod_j psi(b, j, c).

    psi = sigma (phi((l-mu)/sigma) - phi((u-mu)/sigma)) + mu (Phi((u-mu)/sigma) - Phi((l-mu)/sigma));
    infinite bounds give Phi = 0 / 1 and phi = 0.
    """
    means = np.asarray(means, dtype=np.float64)
    std = np.sqrt(np.asarray(variances, dtype=np.float64))
    out = np.zeros(means.shape[0])
    if lower.shape[0] == 0:
        return out
    for i in range(means.shape[0]):
        mu = means[i][None, :]
        sd = std[i][None, :]
        with np.errstate(invalid="ignore"):
            zl = (lower - mu) / sd
            zu = (upper - mu) / sd
        pl = np.where(np.isinf(lower), 0.0, norm.cdf(zl))
        pu = np.where(np.isinf(upper), 1.0, norm.cdf(zu))
        psi = sd * (norm.pdf(zl) - norm.pdf(zu)) + mu * (pu - pl)
        out[i] = np.sum(np.prod(psi, axis=1))
    return out


def select_candidates(front, means, variances, ref, k):
    """hv_box_decomposition.py:306-351: indices of the k largest scores (+ the scores).

    Ties are broken by candidate index (stable), the reference uses numpy's
    default unstable argsort.
    """
    lower, upper = decompose_boxes(front, ref)
    score = batch_ehvi(lower, upper, means, variances)
    sel = np.argsort(-score, kind="stable")[:k]
    return sel, score
