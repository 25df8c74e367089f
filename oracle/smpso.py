"""Oracle: SMPSO swarm updates (row A12 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/SMPSO.py``:
  * ``update_position``  -> SMPSO.py:311-313
  * ``velocity_vector``  -> SMPSO.py:316-348 (scalar r1, r2, w, c1, c2 per call; two random
    leaders from the archive, the one with the larger crowding distance first; constriction
    chi with phi = c1 + c2 if > 4 else 0; clip to +-(xub - xlb) / 2)
  * ``update_strategy``  -> SMPSO.py:187-238 (per-swarm crowding of the swarm's slice of
    y_gen, velocity update, then per-swarm remove_worst of vstack(x_gen[sl], population[sl]))
"""

import numpy as np

from .dda import dda_ens
from .indicators import crowding_distance_metric
from .moea import remove_worst


def update_position(parameters, velocity, xlb, xub):
    """SMPSO.py:311-313."""
    return np.clip(parameters + velocity, xlb, xub)


def constriction(c1, c2):
    """SMPSO.py:322-327: chi = 2 / (2 - phi - sqrt(phi^2 - 4 phi)), phi = c1 + c2 if > 4 else 0."""
    phi = c1 + c2 if c1 + c2 > 4 else 0.0
    return 2.0 / (2.0 - phi - np.sqrt(phi * phi - 4.0 * phi))


def velocity_vector_u(position, velocity, archive, crowding, xlb, xub, u5, leaders):
    """SMPSO.py:316-348 with the five uniform draws ``u5`` in [0,1) and the two leader draws given.

    u5 = (r1, r2, w01, c101, c201): w = 0.1 + 0.4 w01, c = 1.5 + c01.
    """
    r1, r2 = u5[0], u5[1]
    w = 0.1 + (0.5 - 0.1) * u5[2]
    c1 = 1.5 + (2.5 - 1.5) * u5[3]
    c2 = 1.5 + (2.5 - 1.5) * u5[4]
    chi = constriction(c1, c2)
    delta = (np.asarray(xub, dtype=np.float64) - np.asarray(xlb, dtype=np.float64)) / 2
    if archive.shape[0] > 2:
        i1, i2 = int(leaders[0]), int(leaders[1])
        if crowding[i1] < crowding[i2]:
            i1, i2 = i2, i1
    else:
        i1 = i2 = 0
    # archive and position are float32 state arrays in the reference (SMPSO.py:107-113, :184): the
    # difference is formed in their own dtype (float32) and only then promoted by the float64 scalars
    d1 = np.asarray(archive[i1] - position, dtype=np.float64)
    d2 = np.asarray(archive[i2] - position, dtype=np.float64)
    out = (w * np.asarray(velocity, dtype=np.float64) + c1 * r1 * d1 + c2 * r2 * d2) * chi
    return np.clip(out, -delta, delta)


def update_swarm(x_gen_sl, y_gen_sl, pop_x_sl, pop_y_sl, popsize, metric=None, rank_fn=dda_ens):
    """SMPSO.py:218-228 for one swarm: remove_worst(vstack(children, parents))."""
    ym = None if metric is None else [metric]
    return remove_worst(np.vstack((x_gen_sl, pop_x_sl)), np.vstack((y_gen_sl, pop_y_sl)), popsize, ym, rank_fn)


def swarm_crowding(y_gen_sl):
    """SMPSO.py:211-212: crowding of the swarm's offspring objectives."""
    return crowding_distance_metric(y_gen_sl)
