"""Opt-in routing of the reference controller's own helpers to the GPU library.

The optimizer and surrogate plugins are reached by import path and need no patching.  A few hot helpers, however, are
called by the reference's controller code directly on its own modules, so swapping the plugins does not reach them:

  * resample step of ``MOASMO.epoch``     ``MOEA.get_duplicates(best_x, x_0)`` + ``MOEA.crowding_distance_metric``
                                          (dmosopt/MOASMO.py:441-448)
  * ``MOASMO.get_best``                   ``MOEA.get_duplicates(y)`` + ``MOEA.sortMO`` (dmosopt/MOASMO.py:581-639)
  * per-generation termination            ``dmosopt.hv.AdaptiveHyperVolume.compute_hypervolume`` of the whole
                                          population (dmosopt/hv_termination.py:1093-1134 via the multi-fidelity
                                          tracker; dmosopt/hv.py:123-189)
  * the rank function of every ``sortMO`` ``dmosopt.dda.dda_ens`` (dmosopt/dda.py:97-152)

``install()`` rebinds exactly those module attributes of an already importable ``dmosopt`` package to the functions of
this package (same signatures, same results: see tests/test_gpu_reference_loop.py) and ``uninstall()`` restores them.
Nothing is patched implicitly; importing dmosopt_b200 never touches dmosopt.
"""

import importlib

import numpy as np

from . import MOEA as _MOEA
from . import _lib
from . import indicators as _ind

_saved = []


def _set(obj, name, new):
    _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, new)


def _dda_ens(Y, return_dom=False):
    if return_dom:
        raise NotImplementedError("dmosopt_b200: the dense dominance matrix is never materialised (dda.py:40-41 needs O(n^2) memory)")
    return _lib.rank_nd(np.asarray(Y, dtype=np.float64))


def install(package="dmosopt"):
    """Route the helpers listed in the module docstring to the GPU.  Returns the list of patched names."""
    if _saved:
        return [f"{o.__name__}.{n}" for o, n, _ in _saved]
    moea = importlib.import_module(f"{package}.MOEA")
    ind = importlib.import_module(f"{package}.indicators")
    dda = importlib.import_module(f"{package}.dda")
    hv = importlib.import_module(f"{package}.hv")
    for name in ("get_duplicates", "remove_duplicates", "sortMO", "orderMO", "remove_worst"):
        _set(moea, name, getattr(_MOEA, name))
    for mod in (moea, ind):
        for name in ("crowding_distance_metric", "euclidean_distance_metric"):
            if hasattr(mod, name):
                _set(mod, name, getattr(_ind, name))
    _set(dda, "dda_ens", _dda_ens)
    if hasattr(moea, "dda_ens"):
        _set(moea, "dda_ens", _dda_ens)

    original = hv.AdaptiveHyperVolume.compute_hypervolume

    def compute_hypervolume(self, pareto_front, algorithm=None, verbose=False):
        exact = algorithm == "box" or (algorithm in (None, "auto") and self.n_objectives < self.dimension_threshold_exact)
        if exact and 1 <= self.n_objectives <= _lib.HV_MAX_OBJECTIVES:
            pf = np.asarray(pareto_front, dtype=np.float64)
            if len(pf) == 0:
                return 0.0
            return _lib.hypervolume(pf, self.ref_point)  # points not strictly inside ref are ignored, as hv.py:159 does
        return original(self, pareto_front, algorithm, verbose)

    _set(hv.AdaptiveHyperVolume, "compute_hypervolume", compute_hypervolume)
    return [f"{getattr(o, '__name__', o)}.{n}" for o, n, _ in _saved]


def uninstall():
    while _saved:
        obj, name, old = _saved.pop()
        setattr(obj, name, old)
