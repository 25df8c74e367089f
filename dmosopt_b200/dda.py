"""Mirror of ``dmosopt/dda.py``: non-dominated ranking on the GPU (dmo_rank_nd)."""

from . import _lib


def dda_ens(Y, return_dom=False):
    """dda.dda_ens (dda.py:97-133): front index of every row of Y (canonical rank)."""
    if return_dom:
        raise NotImplementedError("the dense dominance-degree matrix is not materialised on the GPU path")
    return _lib.rank_nd(Y)


dda_non_dominated_sort = dda_ens
