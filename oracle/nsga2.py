"""Oracle: NSGA-II plugin steps (rows A9/A10 of SURVEY.md section 8a).

Test infrastructure only (see oracle/__init__.py).

Restates ``dmosopt/NSGA2.py``:
  * ``initialize_state``  -> NSGA2.py:84-114
  * ``generate_strategy`` -> NSGA2.py:116-185 (tournament pool, then the serial
    crossover/mutation loop ``while count < popsize - 1``)
  * ``update_strategy``   -> NSGA2.py:187-236 (children stacked first, remove_worst,
    survivors written back in place -- objectives round to the state dtype)
"""

import numpy as np

from .dda import dda_ens
from .moea import crossover_sbx_u, mutation_u, remove_worst, sort_mo


def initialize(x, y, popsize, metric=None, rank_fn=dda_ens):
    """NSGA2.py:84-114: sortMO then the first ``popsize`` rows."""
    ym = None if metric is None else [metric]
    xs, ys, rank, _, _ = sort_mo(np.asarray(x), np.asarray(y), ym, rank_fn)
    return xs[:popsize], ys[:popsize], rank[:popsize]


def update(pop_x, pop_y, x_gen, y_gen, popsize, metric=None, rank_fn=dda_ens):
    """NSGA2.py:205-214: vstack((x_gen, pop)) -> remove_worst; returns (x, y, rank, perm)."""
    ym = None if metric is None else [metric]
    X = np.vstack((x_gen, pop_x))
    Y = np.vstack((y_gen, pop_y))
    return remove_worst(X, Y, popsize, ym, rank_fn)


def variation_plan(u_cross, u_mut, popsize, crossover_prob=0.9, mutation_prob=0.1):
    """The control flow of NSGA2.py:142-178 for given per-iteration decision draws.

    Iteration t emits 2 children if u_cross[t] < crossover_prob and then 1 more
    if u_mut[t] < mutation_prob; the loop stops at the first t whose starting
    count is >= popsize - 1.  Returns (n_iter, kind, src_iter, slot) per child:
    kind 0 = SBX child 1, 1 = SBX child 2, 2 = mutant.
    """
    kinds, its, count, t = [], [], 0, 0
    while count < popsize - 1:
        if u_cross[t] < crossover_prob:
            kinds += [0, 1]
            its += [t, t]
            count += 2
        if u_mut[t] < mutation_prob:
            kinds += [2]
            its += [t]
            count += 1
        t += 1
    return t, np.asarray(kinds, dtype=np.int32), np.asarray(its, dtype=np.int32)


def generate_given_draws(pool, u_cross, u_mut, pair, single, u_genes, popsize, di_crossover, di_mutation, xlb, xub, mutation_rate,
                         crossover_prob=0.9, mutation_prob=0.1):
    """Offspring of NSGA2.py:142-178 when every random draw is supplied.

    pool      : (poolsize, d) mating pool rows
    pair      : (T, 2) distinct pool indices used by iteration t's crossover
    single    : (T,)  pool index used by iteration t's mutation
    u_genes   : (T, 2, d) gene draws: [t, 0] for the SBX pair, [t, 1] for the mutant
    Returns (x_gen, crossover_indices, mutation_indices).
    """
    n_it, kinds, its = variation_plan(u_cross, u_mut, popsize, crossover_prob, mutation_prob)
    rows, cidx, midx = [], [], []
    for c, (kd, t) in enumerate(zip(kinds, its)):
        if kd == 0:
            c1, c2 = crossover_sbx_u(pool[pair[t, 0]], pool[pair[t, 1]], u_genes[t, 0], di_crossover, xlb, xub)
            rows += [c1, c2]
            cidx += [c, c + 1]
        elif kd == 2:
            rows.append(mutation_u(pool[single[t]], u_genes[t, 1], di_mutation, xlb, xub, mutation_rate))
            midx.append(c)
    return np.vstack(rows), np.asarray(cidx, dtype=int), np.asarray(midx, dtype=int)
