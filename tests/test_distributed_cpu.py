"""world_size-2 gloo tests of the N>1 host plumbing: problem partitioning (no collective) and
the two-all-reduce protocol of the rollout-sharded iteration, against the single-rank oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stomp_motion_planner_icra2011_b200 import _abi, scenes
from stomp_motion_planner_icra2011_b200.distributed import (NumpyShardStandIn, ShardedIteration, partition_problems,
                                                            partition_rollouts)


def test_partition_problems_covers_everything_once():
    for B in (1, 7, 1024, 4096, 4099):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s = partition_problems(B, world, r)
                seen += list(range(B))[s]
            assert seen == list(range(B))
            sizes = [len(range(B)[partition_problems(B, world, r)]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition_rollouts(10, 4, 0)
    assert partition_rollouts(65536, 8, 3) == slice(24576, 32768)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_reference(R=12):
    """one oracle iteration with R rollouts and no reuse: returns the per-rollout arrays and the update."""
    from oracle.oracle import Oracle
    from tests.helpers import correlated_noise
    sc = scenes.make_scenario("tiny", num_problems=1, num_rollouts=R)
    sc.num_reused_rollouts = 0
    o = Oracle(sc, 0)
    rng = np.random.default_rng(5)
    eps = correlated_noise(o.get(_abi.FIELD_NOISE_CHOLESKY), rng, (R,), sc.noise_stddev)
    theta0 = o.get_parameters()
    o.iterate(1, eps)
    return dict(c=o.get(_abi.FIELD_CUMULATIVE_COSTS), eps=o.get(_abi.FIELD_NOISE), M=o.get(_abi.FIELD_PROJECTION),
                theta0=theta0, updates=o.get(_abi.FIELD_UPDATES), theta1=o.get(_abi.FIELD_THETA))


def _worker(rank, world, port, ref, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sl = partition_rollouts(ref["c"].shape[0], world, rank)
    eng = NumpyShardStandIn(ref["c"][sl], ref["eps"][sl], ref["M"], ref["theta0"])
    it = ShardedIteration(eng, eng.minmax, eng.sums, dist=dist)
    it.iterate(1)
    out[rank] = (eng.updates, eng.theta)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_rollout_sharded_iteration_matches_single_rank(world):
    ref = _oracle_reference()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ref, out), nprocs=world, join=True)
    for rank in range(world):
        upd, theta = out[rank]
        np.testing.assert_allclose(upd, ref["updates"], rtol=1e-9, atol=1e-15)
        np.testing.assert_allclose(theta, ref["theta1"], rtol=1e-9, atol=1e-15)
    np.testing.assert_array_equal(out[0][1], out[1][1])       # replicas stay bit-identical


def test_single_rank_stand_in_equals_oracle():
    ref = _oracle_reference()
    eng = NumpyShardStandIn(ref["c"], ref["eps"], ref["M"], ref["theta0"])
    ShardedIteration(eng, eng.minmax, eng.sums).iterate(1)
    np.testing.assert_allclose(eng.updates, ref["updates"], rtol=1e-9, atol=1e-15)
