"""Mirror of the exact-hypervolume entry points on the GPU.

  * AdaptiveHyperVolume.compute_hypervolume   dmosopt/hv.py:123-189 (the 'box' branch, objectives < 10)
  * HyperVolumeBoxDecomposition               dmosopt/hv_box_decomposition.py:62-351
  * compute_hypervolume_box_decomposition     dmosopt/hv_box_decomposition.py:445-464
The Monte-Carlo branches (>= 10 objectives, dmosopt/hv_adaptive.py) are out of scope.
"""

import numpy as np

from . import _lib


class HyperVolumeBoxDecomposition:
    def __init__(self, ref_point):
        self.ref_point = np.asarray(ref_point, dtype=np.float64)
        self.d = len(self.ref_point)

    def compute_hypervolume(self, points):
        points = np.asarray(points, dtype=np.float64)
        if len(points) == 0:
            return 0.0
        if points.shape[1] != self.d:
            raise ValueError(f"Points dimension {points.shape[1]} doesn't match ref point {self.d}")
        return _lib.hypervolume(points, self.ref_point)

    def select_candidates(self, pareto_front, candidate_means, candidate_variances, n_select=1, batch_size=100):
        sel, score = _lib.ehvi_select(pareto_front, candidate_means, candidate_variances, self.ref_point, n_select, nds=False, return_scores=True)
        return sel, score[sel]


def compute_hypervolume_box_decomposition(points, ref_point):
    return HyperVolumeBoxDecomposition(ref_point).compute_hypervolume(points)


class AdaptiveHyperVolume:
    def __init__(self, ref_point, dimension_threshold_exact=10, **kwargs):
        self.ref_point = np.asarray(ref_point, dtype=np.float64)
        self.n_objectives = len(self.ref_point)
        self.dimension_threshold_exact = dimension_threshold_exact

    def compute_hypervolume(self, pareto_front, algorithm=None, verbose=False):
        pareto_front = np.asarray(pareto_front, dtype=np.float64)
        if len(pareto_front) == 0:
            return 0.0
        if algorithm in (None, "auto"):
            algorithm = "box" if self.n_objectives < self.dimension_threshold_exact else "hybrid"
        if algorithm != "box":
            raise NotImplementedError("only the exact 'box' branch (objectives < 10) is accelerated")
        return _lib.hypervolume(pareto_front, self.ref_point)
