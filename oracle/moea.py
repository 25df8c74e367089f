"""Oracle: shared MOEA operators (rows A5-A8, A21 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/MOEA.py``:
  * ``mutation``             -> MOEA.py:191-212  (polynomial mutation, every gene perturbed)
  * ``crossover_sbx``        -> MOEA.py:215-239
  * ``sortMO`` / ``orderMO`` -> MOEA.py:242-347  (rank, optional distances, np.lexsort)
  * ``remove_worst``         -> MOEA.py:398-423
  * ``tournament_selection`` -> MOEA.py:375-395
  * ``get_duplicates``       -> MOEA.py:426-437

The variation operators take the uniform draws ``u`` explicitly so that the CUDA
kernels can be compared gene for gene (the reference draws them from a serial
NumPy generator; only distributional parity is possible for the RNG itself).
"""

import numpy as np

from .dda import dda_ens, rank_canonical
from .indicators import crowding_distance_metric, euclidean_distance_metric


def mutation_u(parent, u, di_mutation, xlb, xub, mutation_rate):
    """MOEA.py:191-212 with the uniform draws ``u`` given.  parent/u: (..., d)."""
    parent = np.asarray(parent, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    di = np.broadcast_to(np.asarray(di_mutation, dtype=np.float64), u.shape)
    lo = u < mutation_rate
    with np.errstate(invalid="ignore"):
        d_lo = (2.0 * u) ** (1.0 / (di + 1)) - 1.0
        d_hi = 1.0 - (2.0 * (1.0 - u)) ** (1.0 / (di + 1))
    delta = np.where(lo, d_lo, d_hi)
    return np.clip(parent + (xub - xlb) * delta, xlb, xub)


def crossover_sbx_u(parent1, parent2, u, di_crossover, xlb, xub):
    """MOEA.py:215-239 with the uniform draws ``u`` given.  Returns (child1, child2)."""
    p1 = np.asarray(parent1, dtype=np.float64)
    p2 = np.asarray(parent2, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    di = np.broadcast_to(np.asarray(di_crossover, dtype=np.float64), u.shape)
    lo = u <= 0.5
    with np.errstate(divide="ignore", invalid="ignore"):
        b_lo = (2.0 * u) ** (1.0 / (di + 1))
        b_hi = (1.0 / (2.0 * (1.0 - u))) ** (1.0 / (di + 1))
    beta = np.where(lo, b_lo, b_hi)
    c1 = np.clip(0.5 * ((1 - beta) * p1 + (1 + beta) * p2), xlb, xub)
    c2 = np.clip(0.5 * ((1 + beta) * p1 + (1 - beta) * p2), xlb, xub)
    return c1, c2


_METRICS = {
    "crowding": crowding_distance_metric,
    "euclidean": euclidean_distance_metric,
}


def order_mo(y, y_distance_metrics=None, rank_fn=dda_ens, x=None, x_distance_metrics=None):
    """MOEA.py:300-347: (perm, rank[perm], dists[perm]).

    ``np.lexsort`` keys are (-x_dists..., -y_dists..., rank): rank ascending is
    the primary key, then each distance descending, stable on original index
    (MOEA.py:285-287).
    """
    y = np.asarray(y)
    rank = rank_fn(y)
    y_dists = []
    for m in y_distance_metrics or []:
        f = m if callable(m) else _METRICS[m]
        y_dists.append(f(y))
    x_dists = []
    for m in x_distance_metrics or []:
        x_dists.append(m(x))
    perm = np.lexsort([-dd for dd in x_dists] + [-dd for dd in y_dists] + [rank])
    return perm, rank[perm], tuple(dd[perm] for dd in y_dists)


def sort_mo(x, y, y_distance_metrics=None, rank_fn=dda_ens):
    """MOEA.py:242-297 with return_perm=True."""
    perm, rank, dists = order_mo(y, y_distance_metrics, rank_fn)
    return x[perm], y[perm], rank, dists, perm


def remove_worst(x, y, pop, y_distance_metrics=None, rank_fn=dda_ens):
    """MOEA.py:398-423: first ``pop`` rows of the sortMO order (+ perm)."""
    xs, ys, rank, _, perm = sort_mo(x, y, y_distance_metrics, rank_fn)
    return xs[:pop], ys[:pop], rank[:pop], perm[:pop]


def tournament_order(*metrics):
    """MOEA.py:388-389: candidates ordered by np.lexsort(metrics) (last metric primary)."""
    return np.lexsort(tuple(np.asarray(m) for m in metrics))


def tournament_probabilities(pop, p=0.5):
    """MOEA.py:375-392: P(i-th best) proportional to p (1-p)^i (normalised).

    Underflows to zero for i > ~1075 in float64, which is why the reference
    raises for pop > ~2150.
    """
    prob = p * (1.0 - p) ** np.arange(pop, dtype=np.float64)
    return prob / prob.sum()


def tournament_selection_reference(rng, pop, poolsize, *metrics):
    """MOEA.py:385-395 verbatim semantics (small pop only)."""
    order = tournament_order(*metrics)
    prob = tournament_probabilities(pop)
    return rng.choice(order, size=poolsize, p=prob, replace=False)


def tournament_selection_gumbel(u, poolsize, *metrics, p=0.5):
    """Log-space restatement that scales to any pop (what the CUDA kernel does).

    Successive sampling without replacement with weights w_i is distributed as
    the top-k of ``log w_i + Gumbel`` keys (Plackett-Luce / Gumbel-top-k), so
    with w_i = p (1-p)^i:  key_i = i*log(1-p) - log(-log(u_i)),  u_i ~ U(0,1),
    where i is the position in the lexsort order and u_i is drawn per position.
    Returns pool indices in draw order (largest key first).
    """
    order = tournament_order(*metrics)
    u = np.asarray(u, dtype=np.float64)
    n = order.shape[0]
    key = np.arange(n, dtype=np.float64) * np.log1p(-p) - np.log(-np.log(u))
    top = np.argsort(-key, kind="stable")[:poolsize]
    return order[top]


def get_duplicates(X, eps=1e-16, Y=None):
    """MOEA.py:426-437: row i of X is a duplicate if some row j<i of Y (default: X itself) has ||x_i-y_j|| <= eps
    (``np.triu_indices(len(X), m=len(Y))`` masks j >= i, diagonal included, also in the two-set form)."""
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    n = X.shape[0]
    dup = np.zeros(n, dtype=bool)
    for i in range(1, n):
        hi = min(i, Y.shape[0])
        if hi == 0:
            continue
        dd = np.sqrt(((Y[:hi] - X[i]) ** 2).sum(axis=1))
        dd[np.isnan(dd)] = np.inf
        dup[i] = np.any(dd <= eps)
    return dup


__all__ = [
    "mutation_u",
    "crossover_sbx_u",
    "order_mo",
    "sort_mo",
    "remove_worst",
    "tournament_order",
    "tournament_probabilities",
    "tournament_selection_reference",
    "tournament_selection_gumbel",
    "get_duplicates",
    "rank_canonical",
]
