"""Multi-GPU sharding of the surrogate generation step.

The path shards by candidates with exactly one exchange step per generation:

  * every rank holds the same optimizer state and the same Philox seed, so ``generate()`` produces the
    same offspring on every rank (replicated, deterministic -- no scatter needed);
  * the GP posterior state is replicated (uploaded once per epoch); rank r predicts only its contiguous
    row block of the offspring (the dense N_pop x N_train contraction is the dominant cost);
  * ONE all-gather of the predicted objectives (P x M float64, < 2 MB at pop 65 536) over NCCL / NVLink;
  * rank / crowding / truncation / hypervolume are global over the merged set and are computed redundantly
    on every rank from the gathered objectives (cheap next to the GP, and it keeps the states identical).

One process per GPU (torchrun); ``torch.distributed`` is plumbing only -- the collective is a plain
all-gather, there is no compute to fuse it with (the consumer is a sort, not a GEMM).
The reference has nothing comparable: its only parallelism is the MPI task farm for *true* objective
evaluations (dmosopt/dmosopt.py:2517-2570), orthogonal to this path.
"""

import numpy as np


def shard_bounds(n_rows, world_size, rank):
    """Contiguous row block [lo, hi) of rank ``rank``; every block has ceil(n/world) rows except the tail."""
    per = -(-n_rows // world_size)
    lo = min(rank * per, n_rows)
    hi = min(lo + per, n_rows)
    return lo, hi, per


class ShardedSurrogate:
    """Wraps a surrogate (``predict`` / ``evaluate``) so that each rank evaluates only its row block.

    Drop-in for ``model.objective`` in MOASMO.optimize (dmosopt/MOASMO.py:110-114): ``evaluate(x)`` returns the
    full (P, M) prediction on every rank.
    """

    def __init__(self, surrogate, group=None, device=None):
        import torch.distributed as dist

        self.sm = surrogate
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        self.device = device
        self.return_mean_variance = getattr(surrogate, "return_mean_variance", False)
        self.nOutput = surrogate.nOutput

    def _gather(self, local, per, n_rows):
        """All-gather equally sized (per, M) blocks and trim to n_rows."""
        import torch

        M = local.shape[1]
        block = np.zeros((per, M), dtype=np.float64)
        block[: local.shape[0]] = local
        if self.world == 1:
            return block[:n_rows]
        if self.backend == "nccl":
            t = torch.from_numpy(block).to(self.device if self.device is not None else "cuda")
            out = torch.empty((self.world * per, M), dtype=torch.float64, device=t.device)
            self.dist.all_gather_into_tensor(out, t, group=self.group)
            return out.cpu().numpy()[:n_rows]
        t = torch.from_numpy(block)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        return torch.cat(outs, dim=0).numpy()[:n_rows]

    def predict(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        n = x.shape[0]
        lo, hi, per = shard_bounds(n, self.world, self.rank)
        if hi > lo:
            mean, var = self.sm.predict(x[lo:hi])
        else:
            mean = np.zeros((0, self.nOutput))
            var = np.zeros((0, self.nOutput))
        both = self._gather(np.hstack((mean, var)), per, n)
        M = self.nOutput
        return np.ascontiguousarray(both[:, :M]), np.ascontiguousarray(both[:, M:])

    def evaluate(self, x):
        if self.return_mean_variance:
            return self.predict(x)
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        n = x.shape[0]
        lo, hi, per = shard_bounds(n, self.world, self.rank)
        local = self.sm.evaluate(x[lo:hi]) if hi > lo else np.zeros((0, self.nOutput))
        return np.ascontiguousarray(self._gather(local, per, n))
