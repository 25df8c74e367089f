"""Stand-alone mirror of the surrogate generation loop, for tests and bench.py.

dmosopt itself drives the plugins from ``MOASMO.optimize`` (dmosopt/MOASMO.py:21-131); that module is
not available on the GPU box, so the same loop (:56-127) is restated here, generator plumbing removed
because a surrogate is always present on this path (MOASMO.py:400-410: the generator never yields then).
"""

from collections import namedtuple

import numpy as np

EpochResults = namedtuple("EpochResults", ["best_x", "best_y", "gen_index", "x", "y", "optimizer"])


def optimize(num_generations, optimizer, model, nInput, nOutput, xlb, xub, popsize=100, initial=None, local_random=None,
             optimize_mean_variance=False, on_generation=None, **kwargs):
    if local_random is None:
        local_random = np.random.default_rng()
    bounds = np.column_stack((xlb, xub))

    def evaluate(x):
        if optimize_mean_variance:
            m, v = model.objective.evaluate(x)
            return np.column_stack((m, np.round(v, 6)))
        return model.objective.evaluate(x)

    x = optimizer.generate_initial(bounds, local_random)
    y = evaluate(x).astype(np.float32)  # MOASMO.py:61-64
    if initial is not None:
        x_initial, y_initial = initial
        if x_initial is not None:
            x = np.vstack((x_initial.astype(np.float32), x))  # MOASMO.py:71-74
        if y_initial is not None:
            y = np.vstack((y_initial.astype(np.float32), y))
    optimizer.initialize_strategy(x, y, bounds, local_random, **kwargs)

    gen_indexes = [np.zeros((x.shape[0],), dtype=np.uint32)]
    x_new, y_new = [], []
    for i in range(1, num_generations + 1):
        x_gen, state_gen = optimizer.generate()  # MOASMO.py:105
        y_gen = evaluate(x_gen)  # MOASMO.py:110-114
        optimizer.update(x_gen, y_gen, state_gen)  # MOASMO.py:116
        x_new.append(x_gen)
        y_new.append(y_gen)
        gen_indexes.append(np.ones((x_gen.shape[0],), dtype=np.uint32) * i)
        if on_generation is not None:
            on_generation(i, optimizer)
    bestx, besty = optimizer.population_objectives
    return EpochResults(bestx, besty, np.concatenate(gen_indexes), np.vstack([x] + x_new), np.vstack([y] + y_new), optimizer)
