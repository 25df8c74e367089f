"""Oracle: exact-GP posterior with a linear prior mean (row A19 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

PARITY UNPINNED.  The reference path is ``EGP_Matern.predict`` (``dmosopt/model_gpytorch.py:2188-2228``): M independent
``gpytorch.models.ExactGP`` (``GPyTorchExactGPModelMatern``, ``:455-508``) with ``LinearMean``, ``ScaleKernel(
MaternKernel(nu=2.5, ard_num_dims=d))`` and ``GaussianLikelihood``; ``predict`` evaluates ``likelihood(model(x))``.
The arithmetic lives in gpytorch 1.13 + linear-operator 0.5.3 (``uv.lock:520-522, 837-839``), which are neither
installed in this image nor vendored in the reference, and the reference holds no test or golden vector for this path.
What follows restates gpytorch's published exact prediction strategy (Rasmussen & Williams eq. 2.25-2.26 with a
non-zero prior mean), in float64:

  x_n  = (x - xlb) / xrng,  xrng = xub - xlb (1 where the range is ~0)          model_gpytorch.py:1965-1967, 1986-1988
  y_n  = (y - mean(y)) / std(y)  (std 0 -> 1)                                    model_gpytorch.py:1992-2009
  m(x) = w . x_n + b                                                             gpytorch.means.LinearMean
  k    = s * (1 + sqrt5 r + 5 r^2 / 3) exp(-sqrt5 r),  r = ||(x - x') / l||     ScaleKernel(MaternKernel(nu=2.5, ARD))
  mean = m(x_*) + k_*^T (K + sigma^2 I)^-1 (y_n - m(X))
  var  = s + sigma^2 - k_*^T (K + sigma^2 I)^-1 k_*        (likelihood(...) adds the noise; exact, i.e. the
                                                            ``fast_pred_var=False`` branch -- LOVE is an approximation of it)
  out  = std(y) * mean + mean(y),  std(y)^2 * var          (float32 arrays)      model_gpytorch.py:2216-2224

The zero-mean special case is cross-checked against oracle/gp.py (which *is* pinned against scikit-learn) in
tests/test_oracle_golden.py.
"""

from dataclasses import dataclass, field
from typing import List

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular

from .gp import MATERN52, kernel_matrix


@dataclass
class EGPObjective:
    alpha: np.ndarray  # (N,)  (K + noise I)^-1 (y_n - m(X))
    L: np.ndarray  # (N,N) lower Cholesky factor of K + noise I
    outputscale: float
    lengthscale: np.ndarray  # (d,)
    noise: float
    weight: np.ndarray  # (d,)
    bias: float
    y_mean: float
    y_std: float


@dataclass
class EGPState:
    X_train: np.ndarray  # (N,d) normalised inputs
    xlb: np.ndarray
    xrng: np.ndarray
    objectives: List[EGPObjective] = field(default_factory=list)


def normalise_y(yin):
    """model_gpytorch.py:1992-2009 (``handle_zeros_in_scale``: a zero standard deviation becomes 1)."""
    yin = np.asarray(yin, dtype=np.float64)
    mean = yin.mean(axis=0)
    std = yin.std(axis=0)
    std = np.where(std == 0.0, 1.0, std)
    return (yin - mean) / std, mean, std


def fit_fixed(xin, yin, xlb, xub, lengthscale, outputscale, noise, weight, bias):
    """Posterior state for given hyper-parameters (one row per objective; training itself is out of scope)."""
    xin = np.asarray(xin, dtype=np.float64)
    yin = np.asarray(yin, dtype=np.float64)
    if yin.ndim == 1:
        yin = yin.reshape(-1, 1)
    xlb = np.asarray(xlb, dtype=np.float64)
    xub = np.asarray(xub, dtype=np.float64)
    xrng = np.where(np.isclose(xub - xlb, 0.0, rtol=1e-6, atol=1e-6), 1.0, xub - xlb)
    xn = (xin - xlb) / xrng
    yn, ymean, ystd = normalise_y(yin)
    M, d = yin.shape[1], xin.shape[1]
    st = EGPState(X_train=xn, xlb=xlb, xrng=xrng)
    for m in range(M):
        ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64)[m] if np.ndim(lengthscale) == 2 else np.asarray(lengthscale, dtype=np.float64), (d,))
        s, nz = float(np.ravel(outputscale)[m]), float(np.ravel(noise)[m])
        w = np.asarray(weight, dtype=np.float64).reshape(M, d)[m]
        b = float(np.ravel(bias)[m])
        K = s * kernel_matrix(xn, xn, ls, MATERN52)
        K[np.diag_indices_from(K)] += nz
        L = cholesky(K, lower=True)
        alpha = cho_solve((L, True), yn[:, m] - (xn @ w + b))
        st.objectives.append(EGPObjective(alpha, L, s, np.array(ls), nz, w, b, float(ymean[m]), float(ystd[m])))
    return st


def predict(st: EGPState, xin):
    """(mean, variance), each (P, M) float32 as the reference returns them."""
    xin = np.asarray(xin, dtype=np.float64)
    if xin.ndim == 1:
        xin = xin.reshape(1, -1)
    xn = (xin - st.xlb) / st.xrng
    P, M = xn.shape[0], len(st.objectives)
    mean = np.empty((P, M))
    var = np.empty((P, M))
    for m, o in enumerate(st.objectives):
        Ks = o.outputscale * kernel_matrix(xn, st.X_train, o.lengthscale, MATERN52)
        mu = xn @ o.weight + o.bias + Ks @ o.alpha
        V = solve_triangular(o.L, Ks.T, lower=True)
        v = np.maximum(0.0, (o.outputscale + o.noise) - np.einsum("ij,ij->j", V, V))
        mean[:, m] = o.y_std * mu + o.y_mean
        var[:, m] = o.y_std**2 * v
    return mean.astype(np.float32), var.astype(np.float32)
