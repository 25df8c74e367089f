"""BASELINE.json configurations other than the bench line, run AT THEIR STATED SIZES through the plugin API.

  C2  ZDT3   d=30 M=2 pop=8192    AGEMOEA + GP N_train=2048
  C3  DTLZ2  d=12 M=3 pop=65536   NSGA2   + GP N_train=4096
  C4  DTLZ7  d=22 M=5 pop=32768   SMPSO (5 swarms) + HV-contribution selection
  C5  WFG4-shaped d=24 M=4 pop=131072 CMAES + GP N_train=4096

One warm-up and two timed generations per configuration (scripts/config_sweep.py), then
  * the invariants of the reference contract (population size, bounds, finiteness);
  * the last update recomputed from its recorded inputs: stored ranks == ranks of the UNROUNDED merged set (exact);
  * a comparison with the CPU oracle on a sub-sample: the rank of sampled points of the merged set satisfies the chain
    identity rank_i = 1 + max{rank_j : j dominates i} evaluated in NumPy against every point, and the surrogate's
    predictions of sampled offspring match the oracle restatement of scikit-learn to 1e-5.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))

from oracle import gp as ogp  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from dmosopt_b200 import _lib

    _lib.context()
    return _lib


def _chain_identity_on_a_sample(Y, rank, n_sample=256, seed=0):
    """CPU check of device ranks at sizes the oracle cannot rank as a whole: for sampled points, the rank is one more than
    the largest rank among the points that dominate them (0 if none) -- the definition dda_ens implements (dda.py:97-152)."""
    Y = np.asarray(Y, dtype=np.float64)
    rng = np.random.default_rng(seed)
    for i in rng.choice(len(Y), size=min(n_sample, len(Y)), replace=False):
        dom = np.all(Y <= Y[i], axis=1) & np.any(Y < Y[i], axis=1)
        assert rank[i] == (rank[dom].max() + 1 if dom.any() else 0), i


def _surrogate_matches_the_oracle(last, n_sample=64):
    sm = last["surrogate"]
    st = ogp.from_sklearn(sm.smlist, sm.xlb, sm.xub)
    idx = np.random.default_rng(1).choice(last["x_gen"].shape[0], size=n_sample, replace=False)
    mean_o, _ = ogp.predict(st, last["x_gen"][idx].astype(np.float64))
    ystd = np.array([o.y_std for o in st.objectives])
    assert np.max(np.abs(last["y_gen"][idx] - mean_o) / np.maximum(np.abs(mean_o), ystd)) < 1e-5


def test_c2_agemoea_pop8192(L):
    import config_sweep as cs
    import dmosopt_b200 as b2

    last = {}
    opt, px, py = cs.run("C2 AGEMOEA", b2.AGEMOEA, 30, 2, 8192, 2048, "zdt3", keep_last=last)
    assert px.shape == (8192, 30) and py.shape == (8192, 2)
    _surrogate_matches_the_oracle(last)
    # AGE-MOEA ranks the de-duplicated merged set (AGEMOEA.py:203-210): survivors carry those ranks
    X = np.vstack((last["before"]["population_parm"], last["x_gen"]))
    Y = np.vstack((last["before"]["population_obj"], last["y_gen"]))
    dup = L.get_duplicates(X)
    r = L.rank_nd(Y[~dup])
    _chain_identity_on_a_sample(Y[~dup], r)
    assert np.array_equal(np.sort(np.asarray(opt.state.rank)), np.sort(r)[:8192])


def test_c3_nsga2_pop65536_d12(L):
    import config_sweep as cs
    import dmosopt_b200 as b2

    last = {}
    opt, px, py = cs.run("C3 NSGA2", b2.NSGA2, 12, 3, 65536, 4096, "dtlz2", keep_last=last, distance_metric=None)
    assert px.shape == (65536, 12)
    _surrogate_matches_the_oracle(last)
    # the last update again, from its recorded inputs, through the stand-alone entry point: identical survivors and ranks
    Xm = np.vstack((last["x_gen"], last["before"]["population_parm"]))
    Ym = np.vstack((last["y_gen"], last["before"]["population_obj"].astype(np.float64)))
    Xo, Yo, rk, perm = L.remove_worst(Xm, Ym, 65536)
    assert np.array_equal(np.asarray(opt.state.rank), rk)  # ranks of the unrounded merged set, exactly
    assert np.array_equal(np.asarray(opt.state.population_obj), Yo.astype(np.float32))
    assert np.array_equal(np.asarray(opt.state.population_parm), Xo)
    full = L.rank_nd(Ym)
    assert np.array_equal(full[perm], rk)
    _chain_identity_on_a_sample(Ym, full)


def test_c4_smpso_pop32768_m5_with_hv_contribution_select(L):
    import config_sweep as cs
    import dmosopt_b200 as b2

    last = {}
    pop, S = 32768, 5
    opt, px, py = cs.run("C4 SMPSO", b2.SMPSO, 22, 5, pop, 4096, "dtlz7", keep_last=last)
    assert last["x_gen"].shape == (2 * S * pop, 22)  # 10 * pop offspring evaluated per generation (SMPSO.py:163-185)
    assert np.array_equal(last["x_gen"], last["x_gen"].astype(np.float32).astype(np.float64))  # the reference's float32 values, clipped (MOEA.py:155)
    assert S * pop - 64 <= px.shape[0] <= S * pop  # de-duplicated population (SMPSO.py:248)
    _surrogate_matches_the_oracle(last)
    # swarm 3 of the last update recomputed from the recorded inputs (the reference's slicing: rows [3 pop, 4 pop) of x_gen)
    sl = slice(3 * pop, 4 * pop)
    Xm = np.vstack((last["x_gen"][sl], last["before"]["population_parm"][sl].astype(np.float64)))
    Ym = np.vstack((last["y_gen"][sl], last["before"]["population_obj"][sl].astype(np.float64)))
    Xo, Yo, rk, perm = L.remove_worst(Xm, Ym, pop)
    assert np.array_equal(np.asarray(opt.state.ranks[3]), rk)
    assert np.array_equal(opt.state.population_obj[sl], Yo.astype(np.float32)) and np.array_equal(opt.state.population_parm[sl], Xo.astype(np.float32))
    _chain_identity_on_a_sample(Ym, L.rank_nd(Ym), n_sample=128)
    # HV-contribution selection (A17) on the result: top-k by score, scores against the oracle on a sub-sample
    from oracle import hv as ohv

    front = py[L.rank_nd(py.astype(np.float64)) == 0].astype(np.float64)[:128]
    mu, var = opt.model.objective.predict(px[:4096])
    ref = py.max(axis=0).astype(np.float64) + 1.0
    sel, score = L.ehvi_select(front, mu, var, ref, 256, return_scores=True)
    assert len(np.unique(sel)) == 256 and np.all(np.isfinite(score))
    assert np.all(score[sel].min() >= np.delete(score, sel).max() - 1e-12)
    sel_o, score_o = ohv.select_candidates(front[L.rank_nd(front) == 0], mu[:256], var[:256], ref, 16)
    np.testing.assert_allclose(score[:256], score_o, rtol=1e-9, atol=1e-300)
    assert last["ms"] < 150.0, last["ms"]  # was 265 ms at a quarter of this size with the per-swarm host loops


def test_c5_cmaes_pop131072_m4(L):
    import config_sweep as cs
    import dmosopt_b200 as b2

    last = {}
    pop = 131072
    opt, px, py = cs.run("C5 CMAES", b2.CMAES, 24, 4, pop, 4096, "dtlz2", keep_last=last)
    assert px.shape == (pop, 24) and py.shape == (pop, 4)
    assert last["x_gen"].shape == (pop // 2, 24)  # mu = pop // 2 offspring per generation (CMAES.py:91, 247)
    _surrogate_matches_the_oracle(last)
    st = opt.state
    assert st.parents_x.shape == (pop, 24) and st.A.shape == (pop, 24, 24) and st.Ainv.shape == (pop, 24, 24) and st.pc.shape == (pop, 24)
    # the selection of the last update: ranks of the merged (offspring + parents) set, chain identity on a sample
    Ym = np.vstack((last["y_gen"], last["before"]["parents_y"]))
    full = L.rank_nd(Ym)
    _chain_identity_on_a_sample(Ym, full, n_sample=128)
    # every stored parent is a row of the merged set and carries that row's rank (which rows are kept follows the
    # reference's order_inv front mapping, CMAES.py:190, pinned at small size by tests/golden/plugins.npz)
    order = np.argsort(Ym[:, 0], kind="stable")
    col0 = Ym[order, 0]
    for i in np.random.default_rng(3).choice(pop, size=512, replace=False):
        row = st.parents_y[i]
        lo, hi = np.searchsorted(col0, row[0], "left"), np.searchsorted(col0, row[0], "right")
        hits = [j for j in order[lo:hi] if np.array_equal(Ym[j], row)]
        assert hits and any(full[j] == st.rank[i] for j in hits), i
    # factors stay consistent: A @ Ainv = I for sampled individuals (updateCholesky keeps the pair in step, CMAES.py:489-537)
    idx = np.random.default_rng(2).choice(pop, size=64, replace=False)
    A, Ainv = np.asarray(st.A[idx]), np.asarray(st.Ainv[idx])
    err = np.abs(np.einsum("nij,njk->nik", A, Ainv) - np.eye(24)).max()
    assert err < 1e-9, err
    assert last["ms"] < 200.0, last["ms"]  # median generation; 1600 ms with the factors on the host, 150 ms with host-side positions / step sizes
