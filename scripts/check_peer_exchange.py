"""torchrun --nproc-per-node W scripts/check_peer_exchange.py : config C3's rollout-sharded iteration with the exchanges done
(a) by NCCL all-reduces between host-synchronised phases and (b) in-kernel over NVLink peer memory (k_shard_stats), on the
same seeds.  Checks: every rank holds the same theta; (a) and (b) agree to rounding (NCCL's summation order is its own); both
agree with an unsharded engine holding all rollouts on rank 0's GPU.  Prints timing of both."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from stomp_motion_planner_icra2011_b200 import _abi, scenes
from stomp_motion_planner_icra2011_b200.distributed import PeerShardedIteration, ShardedIteration, device_views
from stomp_motion_planner_icra2011_b200.engine import Engine


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R_total, iters = int(os.environ.get("R_TOTAL", "4096")), int(os.environ.get("ITERS", "6"))
    sc = scenes.make_scenario("C3", num_problems=1, num_rollouts=R_total // world, seed=7)

    def run(kind):
        eng = Engine(sc, device=local, shard_rank=rank, shard_world=world)
        if kind == "nccl":
            mm, sm = device_views(eng)
            drv = ShardedIteration(eng, mm, sm, dist=dist)
        else:
            drv = PeerShardedIteration(eng, dist)
        drv.iterate(1)
        eng.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for it in range(2, 2 + iters):
            drv.iterate(it)
        eng.synchronize()
        dt = (time.perf_counter() - t0) / iters
        if kind == "peer":
            eng.shard_status()
        theta = eng.get(_abi.FIELD_THETA)[0].copy()
        eng.close()
        return theta, dt

    th_nccl, t_nccl = run("nccl")
    th_peer, t_peer = run("peer")
    # every rank holds the same policy
    for th in (th_nccl, th_peer):
        t = torch.from_numpy(th).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on theta"
    # mapping the peers a second time on the same engines (the epoch restarted while the flags kept the old
    # session's values, so the first waits passed at once): the session restarts cleanly and the result is the same
    def run_reopened():
        eng = Engine(sc, device=local, shard_rank=rank, shard_world=world)
        drv = PeerShardedIteration(eng, dist)
        for it in range(1, 3):
            drv.iterate(it)
        eng.synchronize()
        dist.barrier()
        drv = PeerShardedIteration(eng, dist)      # re-open: flags cleared, epoch 0 again, barrier inside
        for it in range(3, 2 + iters):
            drv.iterate(it)
        eng.synchronize()
        eng.shard_status()
        theta = eng.get(_abi.FIELD_THETA)[0].copy()
        eng.close()
        return theta
    th_re = run_reopened()
    assert np.array_equal(th_re, th_peer), "re-opened peer session differs: %g" % np.abs(th_re - th_peer).max()
    scale = np.abs(th_nccl).max()
    err = np.abs(th_nccl - th_peer).max() / scale
    assert err < 1e-9, err
    if rank == 0:
        full = scenes.make_scenario("C3", num_problems=1, num_rollouts=R_total, seed=7)
        eng = Engine(full, device=local)
        for it in range(1, 2 + iters):
            eng.iterate(it, stats=False)
        th_full = eng.get(_abi.FIELD_THETA)[0]
        err_full = np.abs(th_full - th_peer).max() / scale
        assert err_full < 1e-8, err_full
        print("peer exchange OK: world %d, R %d: nccl %.3f ms/iter, peer %.3f ms/iter; |nccl - peer| %.1e, |unsharded - peer| %.1e"
              % (world, R_total, t_nccl * 1e3, t_peer * 1e3, err, err_full), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
