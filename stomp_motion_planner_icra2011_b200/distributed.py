"""Multi-GPU plumbing: one process per GPU, torch.distributed for the (two tiny) collectives.

Two ways the path shards:

* batched planning (configs C2, C4): independent planning problems are split into contiguous blocks per
  rank; the distance field and robot tables are replicated; there is NO data-path collective.
* one problem with a huge number of rollouts (config C3): rollouts are split over ranks, theta and the
  matrices are replicated, and every iteration needs two all-reduces of [2][D][N] doubles:
      MAX over ranks of {max_r c, max_r -c}      (global min / max of the cumulative cost, needed before the
                                                  exp because the temperature is 10 / (max - min),
                                                  reference: src/policy_improvement.cpp:335-356)
      SUM over ranks of {sum_r e, sum_r e*eps}   (normaliser and probability-weighted noise,
                                                  src/policy_improvement.cpp:353-383)
  after which every rank applies the projection and the update redundantly (identical inputs -> identical
  theta on every rank).  The engine exposes the three phases and the two device buffers; this module does
  the all-reduces on them in place (NCCL on GPUs; the same code runs over gloo with host tensors in the CPU
  tests, where a NumPy stand-in plays the engine).
"""
from __future__ import annotations

import numpy as np


def partition_problems(num_problems: int, world: int, rank: int) -> slice:
    """contiguous block of problems owned by `rank` (first `num_problems % world` ranks get one more)."""
    base, extra = divmod(num_problems, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def partition_rollouts(num_rollouts: int, world: int, rank: int) -> slice:
    if num_rollouts % world:
        raise ValueError("num_rollouts must be divisible by the number of ranks")
    per = num_rollouts // world
    return slice(rank * per, (rank + 1) * per)


class _DeviceArray:
    """__cuda_array_interface__ view of an engine-owned device buffer (no copy), for torch.as_tensor."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def device_views(engine):
    """torch tensors aliasing the engine's minmax / sums device buffers ([2*D*N] float64 each)."""
    import torch
    mm, sm, nbytes = engine.shard_buffers()
    n = nbytes // 8
    dev = torch.device("cuda", engine.desc.device)
    return (torch.as_tensor(_DeviceArray(mm, n), device=dev), torch.as_tensor(_DeviceArray(sm, n), device=dev))


class ShardedIteration:
    """Drives runSingleIteration of ONE problem whose rollouts are sharded over the ranks of `group`.

    `engine` needs: iterate_sharded_phase(iteration, phase), synchronize(); `minmax` / `sums` are tensors that
    alias (or, for the stand-in, are) the engine's reduction buffers.
    """

    def __init__(self, engine, minmax, sums, dist=None, group=None):
        self.engine, self.minmax, self.sums, self.dist, self.group = engine, minmax, sums, dist, group

    def _all_reduce(self, tensor, op):
        if self.dist is None or self.dist.get_world_size(self.group) == 1:
            return
        self.dist.all_reduce(tensor, op=op, group=self.group)

    def iterate(self, iteration_number: int):
        e = self.engine
        e.iterate_sharded_phase(iteration_number, 0)      # rollouts, costs, local {max c, max -c}
        e.synchronize()                                   # engine stream -> the collective's stream
        self._all_reduce(self.minmax, self.dist.ReduceOp.MAX if self.dist else None)
        self._sync_collective()
        e.iterate_sharded_phase(iteration_number, 1)      # local {sum e, sum e*eps}
        e.synchronize()
        self._all_reduce(self.sums, self.dist.ReduceOp.SUM if self.dist else None)
        self._sync_collective()
        e.iterate_sharded_phase(iteration_number, 2)      # projection, theta update, noise-less rollout

    def _sync_collective(self):
        t = self.minmax
        if getattr(t, "is_cuda", False):
            import torch
            torch.cuda.current_stream(t.device).synchronize()


class PeerShardedIteration:
    """The same iteration with the two exchanges done in-kernel over NVLink peer memory (k_shard_stats): no NCCL call and
    no host synchronisation inside an iteration.  torch.distributed is only used once, to all-gather the 64-byte CUDA IPC
    handles of the ranks' exchange buffers."""

    def __init__(self, engine, dist, group=None):
        self.engine = engine
        world = dist.get_world_size(group)
        mine = engine.shard_ipc_handle()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        engine.shard_open_peers(handles)
        dist.barrier(group=group)          # every rank has mapped every buffer before the first store into one

    def iterate(self, iteration_number: int):
        self.engine.iterate_sharded_fused(iteration_number)


class NumpyShardStandIn:
    """CPU stand-in with the engine's three-phase protocol, used by the gloo tests: it holds a shard of
    per-rollout cumulative costs and noise and implements exactly the arithmetic of k_shard_stats' two
    phases and its finalize (kernels.cuh) in NumPy."""

    def __init__(self, cumulative, noise, proj, theta):
        import torch
        self.c, self.eps, self.M, self.theta = cumulative, noise, proj, theta.copy()   # [R_loc][D][N], .., [N][N], [D][N]
        D, N = theta.shape
        self.minmax = torch.zeros(2 * D * N, dtype=torch.float64)
        self.sums = torch.zeros(2 * D * N, dtype=torch.float64)
        self.updates = None

    def iterate_sharded_phase(self, iteration_number, phase):
        D, N = self.theta.shape
        if phase == 0:
            mm = np.concatenate([self.c.max(axis=0).ravel(), (-self.c).max(axis=0).ravel()])
            self.minmax.numpy()[:] = mm
        elif phase == 1:
            mm = self.minmax.numpy()
            mx, mn = mm[:D * N].reshape(D, N), -mm[D * N:].reshape(D, N)
            denom = np.maximum(mx - mn, 1e-8)
            e = np.exp(-10.0 * (self.c - mn) / denom)
            self.sums.numpy()[:] = np.concatenate([e.sum(axis=0).ravel(), (e * self.eps).sum(axis=0).ravel()])
        else:
            s = self.sums.numpy()
            u = s[D * N:].reshape(D, N) / s[:D * N].reshape(D, N)
            self.updates = u @ self.M.T
            self.theta += self.updates

    def synchronize(self):
        pass
