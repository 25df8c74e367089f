"""Host-side logic of the plugins (no GPU): the ``_lib`` calls are replaced by the oracle-backed test seam.

Checks the state handling / dtype flow of dmosopt_b200.NSGA2 and GPR_Matern against golden vectors recorded
from the reference plugin, and -- when the reference checkout is present (build container) -- drives the plugins
from the reference's own, unmodified ``MOASMO.epoch``.
"""

import os
import sys

import numpy as np
import pytest

import fake_backend
from conftest import load_golden

REFERENCE = "/root/reference"


@pytest.fixture
def fake(monkeypatch):
    fake_backend.install(monkeypatch)


def test_nsga2_plugin_reproduces_reference_state_sequence(fake):
    import dmosopt_b200 as b2

    g = load_golden("nsga2")
    for k in range(int(g["ncases"])):
        if bool(g[f"c{k}_ties"]):
            continue  # objective ties: the reference's own result depends on unstable sorts
        metric = str(g[f"c{k}_metric"])
        metric = None if metric == "none" else metric
        x0, y0 = g[f"c{k}_x0"], g[f"c{k}_y0"]
        pop = g[f"c{k}_init_px"].shape[0]
        d, M = x0.shape[1], y0.shape[1]
        bounds = np.column_stack((np.zeros(d), np.ones(d)))
        opt = b2.NSGA2(popsize=pop, nInput=d, nOutput=M, model=b2.Model(), distance_metric=metric)
        opt.initialize_strategy(x0, y0, bounds, np.random.default_rng(1))
        assert np.array_equal(opt.state.population_parm, g[f"c{k}_init_px"])
        assert np.array_equal(opt.state.population_obj, g[f"c{k}_init_py"])
        assert opt.state.population_obj.dtype == np.float32
        assert np.array_equal(opt.state.rank, g[f"c{k}_init_rank"])
        for gi in range(3):
            st = {"crossover_indices": g[f"c{k}_g{gi}_cidx"], "mutation_indices": g[f"c{k}_g{gi}_midx"]}
            opt.update(g[f"c{k}_g{gi}_xgen"], g[f"c{k}_g{gi}_ygen"], st)
            assert np.array_equal(opt.state.population_parm, g[f"c{k}_g{gi}_px"]), (k, gi)
            assert np.array_equal(opt.state.population_obj, g[f"c{k}_g{gi}_py"]), (k, gi)
            assert np.array_equal(opt.state.rank, g[f"c{k}_g{gi}_rank"]), (k, gi)


def test_nsga2_plugin_generate_contract(fake):
    import dmosopt_b200 as b2

    pop, d, M = 50, 7, 2
    rng = np.random.default_rng(0)
    bounds = np.column_stack((-np.ones(d), 2 * np.ones(d)))
    x0 = rng.uniform(-1, 2, size=(pop, d))
    y0 = rng.random((pop, M)).astype(np.float32)
    opt = b2.NSGA2(popsize=pop, nInput=d, nOutput=M, model=b2.Model(), distance_metric=None)
    opt.initialize_strategy(x0, y0, bounds, rng)
    x_gen, st = opt.generate()
    assert x_gen.shape[1] == d and pop - 1 <= x_gen.shape[0] <= pop + 1  # NSGA2.py:142
    assert np.all(x_gen >= bounds[:, 0]) and np.all(x_gen <= bounds[:, 1])
    ci, mi = st["crossover_indices"], st["mutation_indices"]
    assert len(ci) % 2 == 0 and len(ci) + len(mi) == x_gen.shape[0]
    assert sorted(np.concatenate((ci, mi)).tolist()) == list(range(x_gen.shape[0]))
    assert opt.state.total_crossovers == len(ci) // 2 and opt.state.total_mutations == len(mi)
    px, py = opt.population_objectives
    assert px.shape == (pop, d) and py.shape == (pop, M)
    # same seed -> same offspring (Philox seed is drawn from the caller's generator)
    opt2 = b2.NSGA2(popsize=pop, nInput=d, nOutput=M, model=b2.Model(), distance_metric=None)
    opt2.initialize_strategy(x0, y0, bounds, np.random.default_rng(0))
    opt3 = b2.NSGA2(popsize=pop, nInput=d, nOutput=M, model=b2.Model(), distance_metric=None)
    opt3.initialize_strategy(x0, y0, bounds, np.random.default_rng(0))
    assert np.array_equal(opt2.generate()[0], opt3.generate()[0])


def test_gpr_matern_plugin_matches_reference_predictions(fake):
    import dmosopt_b200 as b2

    g = load_golden("gp")
    for k in (0, 1, 5):
        cls = b2.GPR_Matern if str(g[f"c{k}_kind"]) == "matern" else b2.GPR_RBF
        M = g[f"c{k}_yin"].shape[1]
        d = g[f"c{k}_xin"].shape[1]
        sm = cls(g[f"c{k}_xin"], g[f"c{k}_yin"], d, M, g[f"c{k}_xlb"], g[f"c{k}_xub"], optimizer=None)
        mean, var = sm.predict(g[f"c{k}_xtest"])
        np.testing.assert_allclose(mean, g[f"c{k}_mean"], rtol=1e-7, atol=1e-9)
        assert sm.evaluate(g[f"c{k}_xtest"]).shape == mean.shape
        sm.return_mean_variance = True
        assert len(sm.evaluate(g[f"c{k}_xtest"])) == 2


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "dmosopt")), reason="reference checkout not present")
def test_plugins_drop_into_unmodified_moasmo_epoch(fake):
    """The reference's own MOASMO.epoch (dmosopt/MOASMO.py:196-470) drives the plugins by import path."""
    sys.path.insert(0, REFERENCE)
    try:
        from dmosopt import MOASMO
    finally:
        sys.path.remove(REFERENCE)

    d, M, pop = 6, 2, 24
    rng = np.random.default_rng(5)
    xlb, xub = np.zeros(d), np.ones(d)

    def f(x):
        g_ = 1.0 + 9.0 / (d - 1) * x[:, 1:].sum(axis=1)
        return np.column_stack((x[:, 0], g_ * (1.0 - np.sqrt(x[:, 0] / g_))))

    X = rng.random((40, d))
    Y = f(X)
    gen = MOASMO.epoch(
        4, [f"x{i}" for i in range(d)], ["y1", "y2"], xlb, xub, 0.25, X, Y, None, pop=pop,
        optimizer_name="dmosopt_b200.NSGA2", surrogate_method_name="dmosopt_b200.GPR_Matern",
        surrogate_method_kwargs={"anisotropic": False, "optimizer": None}, local_random=rng,
    )
    try:
        next(gen)
        raise AssertionError("epoch should finish without yielding when a surrogate is present")
    except StopIteration as ex:
        res = ex.args[0]
    assert res["x_resample"].shape[1] == d and len(res["x_resample"]) > 0
    assert res["y_pred"].shape[1] == M and res["x_sm"].shape[1] == d and res["y_sm"].shape[1] == M
    assert np.all(res["x_resample"] >= xlb) and np.all(res["x_resample"] <= xub)
    assert type(res["optimizer"]).__module__.startswith("dmosopt_b200")
