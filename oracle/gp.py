"""Oracle: exact-GP posterior mean / variance (row A18 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

The reference path is ``dmosopt/model.py:1254-1275`` (``GPR_Matern.predict`` /
``.evaluate``; ``GPR_RBF`` at :1343-1364), whose arithmetic lives in
scikit-learn (not vendored in the reference; pinned ``scikit-learn 1.5.2`` in
``uv.lock:1613-1614``, 1.9.0 in this image): ``GaussianProcessRegressor.fit``
/ ``.predict`` (``sklearn/gaussian_process/_gpr.py``, Rasmussen & Williams
Alg. 2.1) with kernel ``ConstantKernel * Matern(nu=2.5) + WhiteKernel``
(``model.py:1227-1229``; Matern-5/2 in ``sklearn/gaussian_process/kernels.py``:
``K = (1 + sqrt5 r + 5 r^2 / 3) exp(-sqrt5 r)``, ``r = ||x - x'|| / l``).

  fit      : y_n = (y - mean) / std;  K = c*k(X,X) + (noise + 1e-10) I;
             L = chol(K);  alpha = K^-1 y_n
  predict  : mean = std * (K_* alpha) + mean
             var  = std^2 * max(0, (c + noise) - ||L^-1 K_*^T||^2_col)
  dmosopt  : x is first normalised (x - xlb) / (xub - xlb)  (model.py:1262-1263)
             and the returned variance is std**2 of sklearn's std (model.py:1267)
"""

from dataclasses import dataclass, field
from typing import List

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.spatial.distance import cdist

MATERN52 = 0
RBF = 1
SKLEARN_JITTER = 1e-10  # GaussianProcessRegressor(alpha=1e-10) default, added to diag(K) in fit


def kernel_matrix(X, Y, length_scale, kind=MATERN52):
    """Stationary part of the kernel (without the ConstantKernel factor).

    sklearn ``Matern.__call__`` (nu=2.5) / ``RBF.__call__``; anisotropic when
    ``length_scale`` is a (d,) vector.
    """
    ls = np.asarray(length_scale, dtype=np.float64)
    Xs = np.asarray(X, dtype=np.float64) / ls
    Ys = np.asarray(Y, dtype=np.float64) / ls
    if kind == MATERN52:
        K = cdist(Xs, Ys, metric="euclidean") * np.sqrt(5.0)
        return (1.0 + K + K**2 / 3.0) * np.exp(-K)
    if kind == RBF:
        return np.exp(-0.5 * cdist(Xs, Ys, metric="sqeuclidean"))
    raise ValueError(kind)


@dataclass
class GPObjective:
    """Posterior state of one single-output GP (one entry of ``smlist``)."""

    alpha: np.ndarray  # (N,)  K^-1 y_n
    L: np.ndarray  # (N,N) lower Cholesky factor of K
    constant: float  # ConstantKernel value c
    length_scale: np.ndarray  # scalar or (d,)
    noise: float  # WhiteKernel noise level
    y_mean: float
    y_std: float
    kind: int = MATERN52


@dataclass
class GPState:
    """What ``GPR_Matern`` holds after construction (model.py:1182-1252)."""

    X_train: np.ndarray  # (N,d) already normalised to [0,1]^d
    xlb: np.ndarray
    xub: np.ndarray
    objectives: List[GPObjective] = field(default_factory=list)


def fit_fixed(xin, yin, xlb, xub, constant=1.0, length_scale=0.5, noise=1e-6, kind=MATERN52):
    """``GPR_Matern.__init__`` with fixed hyper-parameters (sklearn ``optimizer=None``).

    model.py:1214-1251: normalise x, one ``GaussianProcessRegressor(normalize_y=True)``
    per objective with the initial theta (c=1, l=0.5, noise=1e-6; RBF noise 1e-5).
    """
    xlb = np.asarray(xlb, dtype=np.float64)
    xub = np.asarray(xub, dtype=np.float64)
    X = (np.asarray(xin, dtype=np.float64) - xlb) / (xub - xlb)
    Y = np.asarray(yin, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    st = GPState(X_train=X, xlb=xlb, xub=xub)
    M = Y.shape[1]
    cs = np.broadcast_to(np.asarray(constant, dtype=np.float64), (M,))
    ns = np.broadcast_to(np.asarray(noise, dtype=np.float64), (M,))
    for m in range(M):
        ls = length_scale[m] if isinstance(length_scale, (list, tuple)) else length_scale
        y = Y[:, m]
        y_mean = float(np.mean(y))
        y_std = float(np.std(y))
        if y_std < 10 * np.finfo(np.float64).eps:  # sklearn _handle_zeros_in_scale
            y_std = 1.0
        yn = (y - y_mean) / y_std
        K = cs[m] * kernel_matrix(X, X, ls, kind)
        K[np.diag_indices_from(K)] += ns[m] + SKLEARN_JITTER
        L = cholesky(K, lower=True, check_finite=False)
        alpha = cho_solve((L, True), yn, check_finite=False)
        st.objectives.append(
            GPObjective(alpha, L, float(cs[m]), np.asarray(ls, dtype=np.float64), float(ns[m]), y_mean, y_std, kind)
        )
    return st


def from_sklearn(smlist, xlb, xub, kind=MATERN52):
    """Extract the posterior state from fitted sklearn regressors (``GPR_Matern.smlist``)."""
    st = GPState(X_train=np.asarray(smlist[0].X_train_, dtype=np.float64), xlb=np.asarray(xlb, float), xub=np.asarray(xub, float))
    for gpr in smlist:
        k = gpr.kernel_
        st.objectives.append(
            GPObjective(
                np.asarray(gpr.alpha_, dtype=np.float64).ravel(),
                np.asarray(gpr.L_, dtype=np.float64),
                float(k.k1.k1.constant_value),
                np.asarray(k.k1.k2.length_scale, dtype=np.float64),
                float(k.k2.noise_level),
                float(np.ravel(gpr._y_train_mean)[0]),
                float(np.ravel(gpr._y_train_std)[0]),
                kind,
            )
        )
    return st


def predict(st: GPState, xin):
    """``GPR_Matern.predict`` (model.py:1254-1268) -> (mean (P,M), var (P,M)), float64."""
    xin = np.asarray(xin, dtype=np.float64)
    if xin.ndim == 1:
        xin = xin[None, :]
    X = (xin - st.xlb) / (st.xub - st.xlb)
    P = X.shape[0]
    M = len(st.objectives)
    mean = np.zeros((P, M))
    var = np.zeros((P, M))
    for m, ob in enumerate(st.objectives):
        Ks = ob.constant * kernel_matrix(X, st.X_train, ob.length_scale, ob.kind)
        mu = Ks @ ob.alpha
        mean[:, m] = ob.y_std * mu + ob.y_mean
        V = solve_triangular(ob.L, Ks.T, lower=True, check_finite=False)
        v = (ob.constant + ob.noise) - np.einsum("ij,ij->j", V, V)
        v[v < 0] = 0.0
        # sklearn returns sqrt(var * std^2); dmosopt squares it back (model.py:1267)
        var[:, m] = np.sqrt(v * ob.y_std**2) ** 2
    return mean, var
