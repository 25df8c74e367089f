"""Mirror of the hot-path part of ``dmosopt/indicators.py`` on the GPU.

  * crowding_distance_metric   indicators.py:12-51
  * euclidean_distance_metric  indicators.py:54-62
  * Hypervolume                indicators.py:213-256
  * HypervolumeImprovement     indicators.py:259-313
(IGD / DistanceIndicator are only used by legacy terminations and are out of scope, SURVEY.md section 2 row 3;
inside the hot path every indicator is built with the identity normalisation, ``zero_to_one=False``.)
"""

import numpy as np

from . import _lib


def crowding_distance_metric(Y):
    return _lib.crowding_distance(Y)


def euclidean_distance_metric(Y):
    return _lib.euclidean_distance(Y)


class _Indicator:
    def __init__(self, ref_point=None, pf=None, nds=False, norm_ref_point=True, ideal=None, nadir=None, zero_to_one=False, **kwargs):
        if zero_to_one:
            raise NotImplementedError("zero_to_one normalisation is not used on the hot path")
        if ref_point is None and pf is not None:
            ref_point = np.atleast_2d(np.asarray(pf)).max(axis=0)
        assert ref_point is not None, "For Hypervolume a reference point needs to be provided!"
        self.ref_point = np.asarray(ref_point, dtype=np.float64)
        self.nds = nds
        self.default_if_empty = 0.0

    def do(self, F, *args, **kwargs):
        F = np.asarray(F)
        if F.ndim == 1:
            F = F[None, :]
        if len(F) == 0:
            return self.default_if_empty
        return self._do(F, *args, **kwargs)


class Hypervolume(_Indicator):
    """indicators.Hypervolume (indicators.py:213-256): HV of F (its rank-0 subset when ``nds``)."""

    def _do(self, F):
        # the GPU hypervolume always reduces to the non-dominated subset first, which does not change the value
        return _lib.hypervolume(F, self.ref_point)


class HypervolumeImprovement(_Indicator):
    """indicators.HypervolumeImprovement (indicators.py:259-313): indices of the k best candidates."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.default_if_empty = []

    def _do(self, F, means, variances, k):
        assert k > 0 and len(F) > 0
        sel = _lib.ehvi_select(F, means, variances, self.ref_point, k, nds=bool(self.nds))
        assert len(sel) > 0
        return np.asarray(sel, dtype=int)
