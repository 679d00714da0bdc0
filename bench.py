#!/usr/bin/env python
"""bench.py — STOMP rollout-timestep evaluations per second on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                    # this engine
    torchrun ... bench.py --gpus N --steps K --warmup W              # N ranks, one per GPU, problem-sharded
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU restatement of the reference

A "step" is one STOMP iteration (PolicyImprovementLoop::runSingleIteration) over the whole batch of
planning problems held by a rank: noise -> rollouts -> cost plugin (FK + SDF + velocity) -> control costs ->
probabilities -> update -> noise-less rollout.  One eval = one (problem, rollout, free timestep) state-cost
evaluation; a steady-state iteration performs B * (R - R_reuse + 1) * N of them (the +1 is the noise-less
rollout, the reused rollouts keep their old state costs exactly like the reference).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stomp_rollout_timestep_evals_per_sec"
UNIT = "evals/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(name, rank, problems):
    from stomp_motion_planner_icra2011_b200 import scenes
    return scenes.make_scenario(name, num_problems=problems, seed=7 + rank)


def evals_per_iteration(sc, B, first):
    rgen = sc.num_rollouts if first else sc.num_rollouts - sc.num_reused_rollouts
    return B * (rgen + 1) * sc.num_time_steps


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (oracle/stomp_oracle.cpp), which tests/test_reference_pinning.py pins to the
# reference's own compiled translation units (oracle/_ref)
# ---------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    name, seed_rank, nprob, first_problem, count, iters = args
    from oracle.oracle import Oracle
    sc = workload(name, seed_rank, nprob)
    total = 0.0
    evals = 0
    for b in range(first_problem, first_problem + count):
        o = Oracle(sc, b)
        o.iterate(1)                          # first iteration (all R rollouts) is warm-up, like the GPU arm
        t0 = time.perf_counter()
        for it in range(2, 2 + iters):
            o.iterate(it)
        total += time.perf_counter() - t0
        evals += iters * evals_per_iteration(sc, 1, False)
        o.close()
    return total, evals


def cpu_baseline(name, budget_s=12.0):
    """single-thread oracle on a bounded sample of the workload (rank 0, N=1)."""
    t, ev = _oracle_worker((name, 0, 4, 0, 1, 20))
    per_iter = t / 20
    iters = max(20, int(budget_s / 4 / per_iter))
    t, ev = _oracle_worker((name, 0, 4, 0, 4, iters))
    return {"value": ev / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first 4 problems of %s, %d iterations each after 1 warm-up iteration (%.1f s), oracle/stomp_oracle.cpp -O2, "
                      "1 thread; ms/iteration/problem = %.3f" % (name, iters, t, 1e3 * t / (4 * iters)),
            "ms_per_iteration_per_problem": 1e3 * t / (4 * iters)}


_REF_STATE = {}


def _ref_step(args):
    """advance problem `p` by `iters` STOMP iterations in this worker process (oracle kept alive per problem)."""
    name, nprob, p, iters = args
    from oracle.oracle import Oracle
    st = _REF_STATE.get(p)
    if st is None:
        sc = _REF_STATE.get("sc")
        if sc is None:
            sc = _REF_STATE["sc"] = workload(name, 0, nprob)
        st = _REF_STATE[p] = {"oracle": Oracle(sc, p), "it": 1}
        st["oracle"].iterate(1)
        st["it"] = 2
    o = st["oracle"]
    t0 = time.perf_counter()
    for _ in range(iters):
        o.iterate(st["it"])
        st["it"] += 1
    return time.perf_counter() - t0, iters * evals_per_iteration(_REF_STATE["sc"], 1, False)


def compiled_reference_rate(name, iters=15):
    """single-thread rate of the reference's OWN translation units (oracle/_ref/libstomp_ref.so, built against the stand-in
    headers of oracle/ref_shim/), when the prebuilt library travelled here.  Reported next to the port, not instead of it:
    the stand-in matrix class evaluates eagerly with temporaries, so this build is slower than the reference would be with
    Eigen 2 and using it as the baseline would flatter the GPU."""
    try:
        from oracle import reference
        if not os.path.exists(reference._LIB_PATH):
            return None
        sc = workload(name, 0, 1)
        sc.movement_duration = float(int((sc.num_time_steps + 11) * sc.discretization))
        ref = reference.ReferenceOptimizer(sc, 0)
        ref.begin()
        ref.iterate(1)
        t0 = time.perf_counter()
        for it in range(2, 2 + iters):
            ref.iterate(it)
        t = time.perf_counter() - t0
        ref.close()
        return {"value": iters * evals_per_iteration(sc, 1, False) / t, "unit": UNIT, "cores": 1, "kind": "reference",
                "ms_per_iteration_per_problem": 1e3 * t / iters,
                "sample": "1 %s problem, %d iterations after 1 warm-up; 13 of the reference's 14 .cpp files compiled unmodified "
                          "against oracle/ref_shim (eager stand-in for Eigen 2: slower than the port, hence not the baseline)"
                          % (name, iters)}
    except Exception as e:   # the library is optional evidence, never a reason to fail the bench
        return {"unavailable": repr(e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    name = args.workload
    t, ev = _oracle_worker((name, 0, 1, 0, 1, 10))
    per_iter = t / 10
    # one step = every worker advances its planning problem by `iters_per_step` iterations; bounded so that the
    # whole steps+warmup run stays within a few minutes
    iters_per_step = max(2, int(min(2.0, 120.0 / max(1, args.steps + args.warmup)) / per_iter))
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        jobs = [(name, procs, p, iters_per_step) for p in range(procs)]
        pool.map(_ref_step, jobs, chunksize=1)          # builds scene + oracle in every worker (untimed)
        for _ in range(args.warmup):
            pool.map(_ref_step, jobs, chunksize=1)
        t0 = time.perf_counter()
        evals = 0
        for _ in range(args.steps):
            res = pool.map(_ref_step, jobs, chunksize=1)
            evals += sum(r[1] for r in res)
        wall = time.perf_counter() - t0
    sc = workload(name, 0, 1)
    value = evals / wall
    sample = ("%d worker processes (host cores: %d), one %s planning problem each, %d STOMP iterations per step; "
              "oracle/stomp_oracle.cpp, g++ -O2 -msse2" % (procs, cores, name, iters_per_step))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": _workload_desc(name, sc, procs), "note": "CPU restatement of the reference (oracle/), pinned to "
                       "the reference's own compiled translation units by tests/test_reference_pinning.py"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    cr = compiled_reference_rate(name)
    if cr is not None:
        line["compiled_reference"] = cr
    print(json.dumps(line))
    return 0


def _workload_desc(name, sc, B):
    return ("%s: %d-DOF arm, N=%d free timesteps, R=%d rollouts/iter (%d reused), K=%d collision spheres, SDF %dx%dx%d u8 @ %.3f m, "
            "%d planning problems per GPU" % (name, sc.robot.num_dimensions, sc.num_time_steps, sc.num_rollouts,
                                              sc.num_reused_rollouts, len(sc.robot.spheres), *sc.sdf.dims, sc.sdf.resolution, B))


# ---------------------------------------------------------------------------------------------------
# this engine
# ---------------------------------------------------------------------------------------------------
DEFAULT_PROBLEMS = {"C1": 1, "C2": 1024, "C4": 512, "C5": 1}


def _load_json(*parts):
    path = os.path.join(ROOT, *parts)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


class Ctx:
    """rank / world / NCCL plumbing shared by the measurements"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        self.json_out = sys.stdout
        if self.world > 1:
            # NCCL writes its version banner to fd 1 when the communicator is created; stdout must carry the JSON line only:
            # keep a private copy of the real stdout for the JSON line and point fd 1 at stderr for everything else
            sys.stdout.flush()
            self.json_out = os.fdopen(os.dup(1), "w")
            os.dup2(2, 1)
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_run(ctx, eng, first_it, K, W):
    """W warm-up iterations, then exactly K iterations (stomp_engine_run) between barriers; device time, max over ranks"""
    eng.run(first_it, W)
    eng.synchronize()
    ctx.barrier()
    l0 = eng.launch_count()
    eng.timer_start()
    eng.run(first_it + W, K)
    ms = eng.timer_stop()
    ctx.barrier()
    return ctx.max_over_ranks(ms), eng.launch_count() - l0, first_it + W + K


def bench_small(ctx, name, dtype, K, W, problems=None):
    """short device-resident measurement of another BASELINE.json config (sub-record of the driver-run line)"""
    from stomp_motion_planner_icra2011_b200.engine import Engine
    B = problems or DEFAULT_PROBLEMS[name]
    sc = workload(name, ctx.rank, B)
    eng = Engine(sc, dtype=dtype, device=ctx.local)
    ms, launches, _ = timed_run(ctx, eng, 1, K, W)
    ev = evals_per_iteration(sc, B, False)
    eng.close()
    return {"workload": _workload_desc(name, sc, B), "ms_per_step": ms / K, "steps": K, "warmup": W, "value": ev * K / (ms * 1e-3), "unit": UNIT,
            "evals_per_step": ev, "gpu_launches": int(launches)}


def bench_c3(ctx, args, dtype):
    """config C3: ONE planning problem, args.rollouts rollouts per iteration sharded over the ranks (strong scaling); per
    iteration two reductions of 2*D*N doubles cross the GPUs — in-kernel over NVLink peer memory (k_shard_stats) and, for
    comparison, as NCCL all-reduces between host-synchronised phases.  Includes a self-check of the sharded policy against an
    unsharded engine (the content of tests/test_gpu_multi.py, which needs more than the one GPU of the test box)."""
    from stomp_motion_planner_icra2011_b200 import scenes, _abi
    from stomp_motion_planner_icra2011_b200.engine import Engine
    from stomp_motion_planner_icra2011_b200.distributed import ShardedIteration, PeerShardedIteration, device_views
    world, rank = ctx.world, ctx.rank
    R_total = args.rollouts
    if R_total % world:
        return {"skipped": "rollouts not divisible by the number of ranks"}
    K, W = max(3, min(args.steps, 10)), 3
    sc = scenes.make_scenario("C3", num_problems=1, num_rollouts=R_total // world, seed=7)
    N, D = sc.num_time_steps, sc.robot.num_dimensions
    evals_step = R_total * N + N          # the noise-less rollout is replicated on every rank: counted once
    out = {"workload": "C3: 1 planning problem, %d rollouts/iteration sharded over %d GPU(s) (%d per GPU), N=%d, %d-DOF, no rollout reuse"
                       % (R_total, world, R_total // world, N, D), "evals_per_step": evals_step, "steps": K, "warmup": W, "unit": UNIT,
           "scaling": "strong"}

    def timed(iterate, first):
        it = first
        for _ in range(W):
            iterate(it); it += 1
        eng.synchronize()
        ctx.barrier()
        eng.timer_start()
        for _ in range(K):
            iterate(it); it += 1
        ms = eng.timer_stop()
        ctx.barrier()
        return ctx.max_over_ranks(ms), it

    eng = Engine(sc, dtype=dtype, device=ctx.local, shard_rank=rank, shard_world=world)
    if world == 1:
        ms, it = timed(lambda i: eng.iterate(i, stats=False), 1)
        out["single_gpu"] = {"ms_per_step": ms / K, "value": evals_step * K / (ms * 1e-3)}
        out["ms_per_step"], out["value"] = ms / K, evals_step * K / (ms * 1e-3)
        out["exchange"] = "none (one GPU): chunked statistics, k_shard_stats with the update fused into the second launch"
    else:
        peer = PeerShardedIteration(eng, ctx.dist)
        ms, it = timed(peer.iterate, 1)
        out["peer"] = {"ms_per_step": ms / K, "value": evals_step * K / (ms * 1e-3),
                       "how": "both exchanges inside k_shard_stats over NVLink peer memory (P2P stores + flags), whole iteration enqueued "
                              "without a host sync, 2 statistics launches per iteration"}
        eng.shard_status()
        mm, sm = device_views(eng)
        nccl = ShardedIteration(eng, mm, sm, dist=ctx.dist)
        ms2, it = timed(nccl.iterate, it)
        out["nccl"] = {"ms_per_step": ms2 / K, "value": evals_step * K / (ms2 * 1e-3),
                       "how": "two NCCL all-reduces of 2*D*N doubles between three host-synchronised phases"}
        out["ms_per_step"], out["value"] = ms / K, evals_step * K / (ms * 1e-3)
        out["exchange"] = "peer"
    eng.close()
    # self-check: 3 sharded iterations against an unsharded engine holding all rollouts (Philox streams are keyed by the global
    # rollout id, so the draw does not depend on the sharding); every rank must hold the same policy
    chk = Engine(sc, dtype=dtype, device=ctx.local, shard_rank=rank, shard_world=world)
    drv = PeerShardedIteration(chk, ctx.dist) if world > 1 else None
    for i in (1, 2, 3):
        drv.iterate(i) if drv else chk.iterate(i, stats=False)
    th = chk.get_parameters()
    chk.close()
    diff_ranks = 0.0
    if world > 1:
        t = ctx.torch.tensor(th, device="cuda")
        t0 = t.clone()
        ctx.dist.broadcast(t0, src=0)
        diff_ranks = ctx.max_over_ranks(float((t - t0).abs().max().item()))
    if rank == 0:
        sc1 = scenes.make_scenario("C3", num_problems=1, num_rollouts=R_total, seed=7)
        one = Engine(sc1, dtype=dtype, device=ctx.local)
        for i in (1, 2, 3):
            one.iterate(i, stats=False)
        th1 = one.get_parameters()
        one.close()
        d = float(np.abs(th - th1).max())
        out["self_check"] = {"iterations": 3, "max_abs_theta_diff_vs_unsharded": d, "theta_matches_unsharded": bool(d <= 1e-12),
                             "max_abs_theta_diff_between_ranks": diff_ranks, "ranks_bit_identical": bool(diff_ranks == 0.0)}
    ctx.barrier()
    return out


def run_engine(args):
    from stomp_motion_planner_icra2011_b200 import _abi
    from stomp_motion_planner_icra2011_b200.engine import Engine
    ctx = Ctx()
    torch, rank, world, local = ctx.torch, ctx.rank, ctx.world, ctx.local
    name = args.workload
    dtype = _abi.F32 if args.dtype == "f32" else _abi.F64
    if name == "C3":      # stand-alone C3 run: the sharded record is the line
        c3 = bench_c3(ctx, args, dtype)
        if rank == 0:
            line = {"metric": METRIC, "value": c3["value"], "unit": UNIT, "n_gpus": world, "steps": c3["steps"], "warmup": c3["warmup"],
                    "ms_per_step": c3["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": args.dtype, "data": "synthetic", "config": {"workload": c3["workload"]}, "c3": c3}
            ctx.json_out.write(json.dumps(line) + "\n")
            ctx.json_out.flush()
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return 0

    B = args.problems or DEFAULT_PROBLEMS.get(name, 1024)
    sc = workload(name, rank, B)
    eng = Engine(sc, dtype=dtype, device=local)
    D, N, R = eng.D, eng.N, eng.R
    rgen = R - sc.num_reused_rollouts
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident throughput ("value"): inputs resident in HBM, engine Philox noise --------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, it = timed_run(ctx, eng, 1, K, W)
    clocks = sampler.stop() if rank == 0 else None
    evals_step = evals_per_iteration(sc, B, False)        # per rank
    total_evals_step = world * evals_step
    value = total_evals_step * K / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ("e2e") ----------------------------------------------------------
    # What PolicyImprovementLoop::runSingleIteration(iteration) takes and returns: the iteration number in (the noise scales
    # sigma_d decay_d^(it-1), D doubles, are uploaded by stomp_engine_iterate), and every step the updated trajectories, the
    # noise-less rollout's cost and its collision flag out to pinned host memory.  Results of step i are requested
    # asynchronously and collected while step i+1 runs (two host buffer sets).
    Ke = max(3, min(K, 20))
    res = [(torch.empty((B, D, N), dtype=torch.float64).pin_memory().numpy(),
            torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
            torch.empty(B, dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]
    d2h = int(res[0][0].nbytes + B * 12)
    pending = []

    def drain():
        while pending:
            eng.wait_results(pending.pop(0))

    def prod_step(i):
        eng.iterate(i, stats=False)
        th, co, cf_ = res[i % 2]
        pending.append(eng.request_results_async(th, co, cf_))
        if len(pending) > 1:
            eng.wait_results(pending.pop(0))

    def time_steps(step, first):
        i = first
        for _ in range(2):
            step(i); i += 1
        drain()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            step(i); i += 1
        drain()                                     # every step's results are in host memory when the clock stops
        torch.cuda.synchronize()
        return ctx.max_over_ranks(time.perf_counter() - t0), i

    prod_s, it = time_steps(prod_step, it)
    e2e = {"value": total_evals_step * Ke / prod_s, "unit": UNIT, "h2d_bytes_per_step": 8 * D, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * prod_s / Ke, "steps": Ke,
           "api": "per step: stomp_engine_iterate(iteration) [uploads the D noise scales; noise is the engine's Philox stream, like the "
                  "reference's runSingleIteration(iteration) which takes no host data either] + stomp_engine_request_results_async("
                  "pinned theta [B][D][N], noise-less cost [B], collision flag [B]) + stomp_engine_wait_results(previous step)"}
    # the parity mode for comparison: the host ALSO supplies every step's exploration noise (B*R_gen*D*N doubles from pinned
    # memory).  At 28.7 MB per 0.5 ms step this is a 57 GB/s stream per GPU: one PCIe Gen5 x16 link each, and 8 ranks read
    # ~460 GB/s from one host NUMA node, which is why this mode cannot scale across the box.
    L = np.linalg.cholesky(eng.get(_abi.FIELD_INV_CONTROL_COST))
    rng = np.random.default_rng(1234 + rank)
    eps_pinned = torch.empty((B, rgen, D, N), dtype=torch.float64).pin_memory()
    eps_np = eps_pinned.numpy()
    z = rng.standard_normal((min(B, 16), rgen, D, N))
    eps_np[...] = np.resize(np.einsum("ij,...j->...i", L, z) * 2.0, eps_np.shape)
    eng.inject_noise_async(eps_np)              # noise of the first step

    def inj_step(i):
        eng.iterate(i, stats=False)             # consumes the pending injection (device-side wait on the copy)
        th, co, cf_ = res[i % 2]
        pending.append(eng.request_results_async(th, co, cf_))
        eng.inject_noise_async(eps_np)          # next step's noise: H2D on the copy stream, overlaps this iteration
        if len(pending) > 1:
            eng.wait_results(pending.pop(0))

    inj_s, it = time_steps(inj_step, it)
    eng.iterate(it, stats=False); it += 1        # consumes the last pending injection
    eng.synchronize()
    e2e_injected = {"value": total_evals_step * Ke / inj_s, "unit": UNIT, "h2d_bytes_per_step": int(eps_np.nbytes + 8 * D),
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * inj_s / Ke, "steps": Ke,
                    "host_read_gbs_all_ranks": world * eps_np.nbytes / (inj_s / Ke) / 1e9,
                    "note": "host-injection (parity) mode: + stomp_engine_inject_noise_async(pinned eps of the next step) every step; "
                            "bound by host memory / PCIe, not by the GPUs, when several ranks share one host"}

    # ---- per-kernel event timing for the roofline (separate pass: events around every launch, one stream) ------------------
    eng.set_profiling(1)
    Kp = max(3, min(K, 20))
    for _ in range(Kp):
        eng.iterate(it, stats=False)
        it += 1
    cost_ms, cost_n = eng.get_profile("k_cost")
    all_ms, all_n = eng.get_profile("")
    shares = {}
    for kn in ("k_select_reuse", "k_generate", "k_gather_state", "k_cost", "k_cumulative", "k_totals", "k_update", "k_extra_total",
               "k_shard_stats", "k_finalize"):
        m, n = eng.get_profile(kn)
        if n:
            shares[kn] = {"ms_per_iteration": m / Kp, "launches_per_iteration": n / Kp}
    eng.set_profiling(0)
    peaks, peak_kind = _peaks()
    mine = _load_json("profiles", "r2_peaks.json") or {}
    counts = (_load_json("profiles", "r2_kcost_counts.json") or {}) if (name == "C2" and B == 1024 and args.dtype == "f64") else {}
    # algorithmic bytes per iteration of the streaming kernels: what has to cross HBM once
    BRDN8 = B * R * D * N * 8.0
    algo = {"k_generate": 4.0 * BRDN8,                      # per vector: read theta / previous parameters, write noise, parameters, control costs
            "k_cumulative": BRDN8 * (2.0 + 1.0 / D),        # read control costs + state costs, write cumulative costs
            "k_totals": B * R * (N + D) * 8.0,              # read state costs + per-vector control-cost sums
            "k_update": 2.0 * BRDN8}                        # read cumulative (or control + state) costs + noise
    for kn, nbytes in algo.items():
        if kn in shares and shares[kn]["ms_per_iteration"] > 0:
            gbs = nbytes / (shares[kn]["ms_per_iteration"] * 1e-3) / 1e9
            shares[kn].update({"algorithmic_bytes_per_iteration": nbytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
    Ksph = len(sc.robot.spheres)
    vox_bytes = {0: 4, 1: 1, 2: 2}[sc.sdf.voxel_dtype]
    bytes_per_eval = D * 8 + Ksph * vox_bytes + 8 + 1
    evals_launch = evals_step / 2.0                                   # two k_cost launches per iteration (new rollouts, noise-less)
    avg_launch_s = cost_ms * 1e-3 / max(cost_n, 1)
    achieved = bytes_per_eval * evals_launch / avg_launch_s / 1e9
    # the roofs k_cost could be bound by, each = (per-eval count of the committed ncu capture of this build) x (evals per launch)
    # / (this run's event-timed launch duration) against a peak measured on this pool (scripts/peaks.cu)
    roofs = {"hbm": {"achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "per_eval": bytes_per_eval, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)"}}
    if counts and mine:
        def roof(per_eval, peak, unit, what):
            rate = per_eval * evals_launch / avg_launch_s
            return {"achieved": rate, "peak": peak, "unit": unit, "frac": rate / peak, "per_eval": per_eval, "what": what}
        roofs["issue"] = roof(counts["warp_inst_per_eval"], mine["issue_warp_inst_per_s"], "warp instructions/s",
                              "all warp instructions (ncu smsp__inst_executed.sum) against the measured FFMA issue rate of the chip")
        roofs["fp64"] = roof(counts["fp64_warp_inst_per_eval"] * 64.0, mine["fp64_fma_tflops"] * 1e12, "flop/s (fp64 pipe, 1 warp instruction = 64 flop)",
                             "fp64-pipe warp instructions (ncu smsp__inst_executed_pipe_fp64.sum) against the measured DFMA rate")
        roofs["l2_gather"] = roof(counts["l2_tex_read_sectors_per_eval"] * 32.0, mine["l2_gather_sector_gbs"] * 1e9, "B/s of 32-byte L2 sectors",
                                  "L2 read sectors requested by the SMs (ncu lts__t_sectors_srcunit_tex_op_read.sum) against the measured random-gather rate")
        binding = max((k for k in roofs), key=lambda k: roofs[k]["frac"])
    else:
        binding = "hbm"
    roofline = {"bound": "hbm", "kernel": "k_cost", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": counts.get("dram_bytes_per_launch"),
                "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "algorithmic_bytes_per_eval": bytes_per_eval, "avg_launch_ms": 1e3 * avg_launch_s,
                "kernel_share_of_step": cost_ms / all_ms if all_ms else None,
                "roofs": roofs, "nearest_roof": binding,
                "counts_source": counts.get("source"), "peaks_source": mine.get("how"),
                "note": "k_cost is a latency-bound fp64 / integer kernel over an L2-resident u8 grid: no roof is near (the nearest is "
                        "instruction issue); HBM is the contract's `bound`, `roofs` holds the measured alternatives",
                "kernels": shares}

    extras = {}
    if name == "C2" and not args.no_extras:
        # the other BASELINE.json configurations, briefly, so that their numbers are driver-run too
        if world == 1:
            for other, (k_, w_) in (("C1", (48, 8)), ("C5", (10, 3)), ("C4", (5, 3))):
                try:
                    extras[other] = bench_small(ctx, other, dtype, k_, w_)
                except Exception as ex:      # a sub-record never fails the headline
                    extras[other] = {"error": repr(ex)}
            if args.dtype == "f64":
                try:      # the fp32 instantiation of the cost plugin on the headline workload (north_star's 1e-3 mode)
                    extras["C2_f32"] = bench_small(ctx, "C2", _abi.F32, 10, 3)
                    extras["C2_f32"]["note"] = "dtype = STOMP_F32: FK, sphere transforms, potentials and velocities in fp32; PI^2 statistics stay fp64"
                except Exception as ex:
                    extras["C2_f32"] = {"error": repr(ex)}
        try:
            extras["C3"] = bench_c3(ctx, args, dtype)
        except Exception as ex:
            extras["C3"] = {"error": repr(ex)}
            ctx.barrier()

    if rank == 0:
        cpu = cpu_baseline(name) if world == 1 and not args.no_cpu_baseline else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic",
                "config": {"workload": _workload_desc(name, sc, B),
                           "parallelism": "problems sharded over %d GPU(s), no collective" % world,
                           "l2": "per-iteration working set %.2f GB of rollout arrays >> 126 MB L2 (no flush needed)"
                                 % (6 * B * R * D * N * 8 / 1e9),
                           "noise": "engine Philox RNG", "evals_per_step_per_gpu": evals_step,
                           "evals_per_step_per_gpu_without_noiseless_rollout": B * rgen * N,
                           "value_without_noiseless_rollout": world * B * rgen * N * K / (ms * 1e-3),
                           "iterations_per_sec": 1e3 * K / ms,
                           "problem_iterations_per_sec": world * B * 1e3 * K / ms,
                           "rollouts_per_sec": total_evals_step / N * 1e3 * K / ms},
                "clocks": clocks, "e2e": e2e, "e2e_injected": e2e_injected, "gpu_launches": int(launches), "roofline": roofline}
        if "C3" in extras:
            line["c3"] = extras.pop("C3")
        if extras:
            line["other_workloads"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
        ctx.json_out.write(json.dumps(line) + "\n")
        ctx.json_out.flush()
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--problems", type=int, default=None,
                    help="planning problems per GPU (default: what BASELINE.json names: C1 1, C2 1024, C4 4096 / 8, C5 1)")
    ap.add_argument("--rollouts", type=int, default=65536, help="total rollouts of the C3 workload")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1 / C4 / C5 / C3 sub-records of the C2 line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_engine(args)


if __name__ == "__main__":
    sys.exit(main())
