"""World-size-2 gloo test of the multi-GPU plumbing (no GPU): candidate sharding + one all-gather."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    import fake_backend
    from dmosopt_b200 import _lib

    for name in fake_backend.FUNCTIONS:  # oracle-backed seam instead of the CUDA library (CPU-only container)
        setattr(_lib, name, getattr(fake_backend, name))
    import dmosopt_b200 as b2
    from dmosopt_b200.driver import optimize
    from dmosopt_b200.parallel import ShardedSurrogate, shard_bounds

    rng = np.random.default_rng(0)
    d, M, N, pop = 5, 2, 60, 31
    xlb, xub = np.zeros(d), np.ones(d)
    Xtr = rng.random((N, d))
    Ytr = np.column_stack((Xtr[:, 0] + Xtr[:, 1] ** 2, 1 - Xtr[:, 0] + Xtr[:, 2]))
    sm = b2.GPR_Matern(Xtr, Ytr, d, M, xlb, xub, optimizer=None)
    sh = ShardedSurrogate(sm)
    X = rng.random((47, d))  # not divisible by the world size
    mean_full, var_full = sm.predict(X)
    mean_sh, var_sh = sh.predict(X)
    assert np.array_equal(mean_full, mean_sh) and np.array_equal(var_full, var_sh)
    assert np.array_equal(sh.evaluate(X), mean_full)
    lo, hi, per = shard_bounds(47, world, rank)
    assert per == 24 and (lo, hi) == ((0, 24) if rank == 0 else (24, 47))
    # replicated optimizer + sharded surrogate: every rank reaches the same population
    mdl = b2.Model(objective=sh)
    opt = b2.NSGA2(popsize=pop, nInput=d, nOutput=M, model=mdl, distance_metric=None)
    res = optimize(3, opt, mdl, d, M, xlb, xub, popsize=pop, local_random=np.random.default_rng(7))
    np.save(os.path.join(out_dir, f"best_y_{rank}.npy"), res.best_y)
    np.save(os.path.join(out_dir, f"best_x_{rank}.npy"), res.best_x)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_surrogate_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    y0, y1 = np.load(tmp_path / "best_y_0.npy"), np.load(tmp_path / "best_y_1.npy")
    x0, x1 = np.load(tmp_path / "best_x_0.npy"), np.load(tmp_path / "best_x_1.npy")
    assert np.array_equal(y0, y1) and np.array_equal(x0, x1)
    assert y0.shape == (31, 2)


def test_shard_bounds_cover_everything():
    from dmosopt_b200.parallel import shard_bounds

    for n in (1, 7, 64, 65537):
        for w in (1, 2, 4, 8):
            rows = []
            for r in range(w):
                lo, hi, per = shard_bounds(n, w, r)
                assert hi - lo <= per
                rows.extend(range(lo, hi))
            assert rows == list(range(n))
