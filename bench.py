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
def run_engine(args):
    import torch
    from stomp_motion_planner_icra2011_b200 import _abi
    from stomp_motion_planner_icra2011_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    json_out = sys.stdout
    if world > 1:
        # NCCL writes its version banner to fd 1 when the communicator is created; stdout must carry the JSON line only:
        # keep a private copy of the real stdout for the JSON line and point fd 1 at stderr for everything else
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    name = args.workload
    dtype = _abi.F32 if args.dtype == "f32" else _abi.F64
    sharded = name == "C3"
    if sharded:
        # one planning problem, rollouts sharded over the ranks (strong scaling), two NCCL all-reduces per iteration
        from stomp_motion_planner_icra2011_b200 import scenes
        from stomp_motion_planner_icra2011_b200.distributed import ShardedIteration, device_views
        B = 1
        total_rollouts = args.rollouts
        sc = scenes.make_scenario("C3", num_problems=1, num_rollouts=total_rollouts // world, seed=7)
        eng = Engine(sc, dtype=dtype, device=local, shard_rank=rank, shard_world=world)
        if world > 1:
            if args.c3_exchange == "peer":   # both exchanges in-kernel over NVLink peer memory (k_peer_allreduce)
                from stomp_motion_planner_icra2011_b200.distributed import PeerShardedIteration
                driver = PeerShardedIteration(eng, dist)
            else:                            # two NCCL all-reduces between host-synchronised phases
                mm, sm = device_views(eng)
                driver = ShardedIteration(eng, mm, sm, dist=dist)
            eng_iterate = lambda i, stats=False: driver.iterate(i)
        else:
            eng_iterate = lambda i, stats=False: eng.iterate(i, stats=stats)
    else:
        B = args.problems or {"C1": 1, "C2": 1024, "C4": 512, "C5": 1}.get(name, 1024)
        sc = workload(name, rank, B)
        eng = Engine(sc, dtype=dtype, device=local)
        eng_iterate = lambda i, stats=False: eng.iterate(i, stats=stats)
    D, N, R = eng.D, eng.N, eng.R
    rgen = R - sc.num_reused_rollouts
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident throughput ("value") -----------------------------------------------------
    it = 1
    for _ in range(W):
        eng_iterate(it)
        it += 1
    eng.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = eng.launch_count()
    eng.timer_start()
    for _ in range(K):
        eng_iterate(it)
        it += 1
    ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - l0
    ms = max_over_ranks(ms)
    clocks = sampler.stop() if rank == 0 else None
    evals_step = evals_per_iteration(sc, B, False)        # per rank
    if sharded:   # the noise-less rollout is replicated on every rank: count it once
        total_evals_step = world * B * rgen * N + N
    else:
        total_evals_step = world * evals_step
    value = total_evals_step * K / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ("e2e") ----------------------------------------
    # host-injection mode: every step copies that step's noise from pinned host memory, runs the iteration,
    # and reads the per-problem noise-less cost / collision flag and the updated trajectories back.
    theta_pinned = torch.empty((B, D, N), dtype=torch.float64).pin_memory()
    theta_np = theta_pinned.numpy()
    Ke = max(3, min(K, 20))
    if not sharded:
        L = np.linalg.cholesky(eng.get(_abi.FIELD_INV_CONTROL_COST))
        rng = np.random.default_rng(1234 + rank)
        eps_pinned = torch.empty((B, rgen, D, N), dtype=torch.float64).pin_memory()
        eps_np = eps_pinned.numpy()
        z = rng.standard_normal((min(B, 16), rgen, D, N))
        base = np.einsum("ij,...j->...i", L, z) * 2.0
        eps_np[...] = np.resize(base, eps_np.shape)

        # results of step i are requested asynchronously and collected while step i+1 already runs (two host buffer sets)
        res = [(torch.empty((B, D, N), dtype=torch.float64).pin_memory().numpy(),
                torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
                torch.empty(B, dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]
        pending = []
        eng.inject_noise_async(eps_np)              # noise of the first timed step

        def e2e_step(i):
            eng.iterate(i, stats=False)             # consumes the pending injection (device-side wait on the copy)
            th, co, cf_ = res[i % 2]
            pending.append(eng.request_results_async(th, co, cf_))   # D2H of theta / noise-less cost / collision flag
            eng.inject_noise_async(eps_np)          # next step's noise: H2D on the copy stream, overlaps this iteration
            if len(pending) > 1:
                eng.wait_results(pending.pop(0))    # step i-1's results are now in host memory

        def e2e_drain():
            while pending:
                eng.wait_results(pending.pop(0))
        h2d, api = int(eps_np.nbytes + 8 * D), ("per step: stomp_engine_iterate + stomp_engine_request_results_async(pinned theta, cost, flag) + "
                                                "stomp_engine_inject_noise_async(pinned eps of the next step) + "
                                                "stomp_engine_wait_results(previous step)")
    else:
        def e2e_step(i):
            eng_iterate(i)
            eng.get_parameters(theta_np)

        def e2e_drain():
            pass
        h2d, api = 8 * D, "sharded iterate (3 phases + 2 all-reduces) + stomp_engine_get_parameters(pinned); noise is engine Philox"
    for _ in range(2):
        e2e_step(it); it += 1
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step(it)
        it += 1
    e2e_drain()                                     # every step's results are in host memory when the clock stops
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    # the production mode for comparison: the engine draws its own noise (the reference's runSingleIteration(iteration) takes no
    # host input either), every step's trajectories / cost / flag still come back to pinned host memory
    rng_mode = None
    if not sharded:
        pend2 = []

        def rng_step(i):
            eng.iterate(i, stats=False)
            th, co, cf_ = res[i % 2]
            pend2.append(eng.request_results_async(th, co, cf_))
            if len(pend2) > 1:
                eng.wait_results(pend2.pop(0))
        for _ in range(2):
            rng_step(it); it += 1
        while pend2:
            eng.wait_results(pend2.pop(0))
        barrier()
        t1 = time.perf_counter()
        for _ in range(Ke):
            rng_step(it); it += 1
        while pend2:
            eng.wait_results(pend2.pop(0))
        torch.cuda.synchronize()
        rng_s = max_over_ranks(time.perf_counter() - t1)
        rng_mode = {"value": total_evals_step * Ke / rng_s, "unit": UNIT, "ms_per_step": 1e3 * rng_s / Ke, "h2d_bytes_per_step": 8 * D,
                    "d2h_bytes_per_step": int(theta_np.nbytes + B * 12),
                    "note": "engine Philox noise instead of host-injected noise; results still read back every step"}
    e2e = {"value": total_evals_step * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(theta_np.nbytes + B * 12), "ms_per_step": 1e3 * e2e_s / Ke, "steps": Ke, "api": api}
    if rng_mode is not None:
        e2e["engine_noise_mode"] = rng_mode

    # ---- per-kernel event timing for the roofline (separate pass: events around every launch) ---------
    eng.set_profiling(1)
    Kp = max(3, min(K, 20))
    for _ in range(Kp):
        eng_iterate(it)
        it += 1
    cost_ms, cost_n = eng.get_profile("k_cost")
    all_ms, all_n = eng.get_profile("")
    shares = {}
    for kn in ("k_select_reuse", "k_generate", "k_gather_state", "k_cost", "k_cumulative", "k_update", "k_extra_total",
               "k_minmax_partial", "k_sums_partial", "k_pair_reduce", "k_finalize"):
        m, n = eng.get_profile(kn)
        shares[kn] = {"ms_per_iteration": m / Kp, "launches_per_iteration": n / Kp}
    eng.set_profiling(0)
    peaks, peak_kind = _peaks()
    # algorithmic bytes per iteration of the streaming kernels: what has to cross HBM once
    BRDN8 = B * R * D * N * 8.0
    algo = {"k_generate": 4.0 * BRDN8,                      # per vector: read theta / previous parameters, write noise, parameters, control costs
            "k_cumulative": BRDN8 * (2.0 + 1.0 / D),        # read control costs + state costs, write cumulative costs
            "k_update": 2.0 * BRDN8}                        # read cumulative costs + noise
    for kn, nbytes in algo.items():
        ms_it = shares[kn]["ms_per_iteration"]
        if ms_it > 0:
            gbs = nbytes / (ms_it * 1e-3) / 1e9
            shares[kn].update({"algorithmic_bytes_per_iteration": nbytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
    Ksph = len(sc.robot.spheres)
    vox_bytes = {0: 4, 1: 1, 2: 2}[sc.sdf.voxel_dtype]
    bytes_per_eval = D * 8 + Ksph * vox_bytes + 8 + 1
    algo_bytes_launch = bytes_per_eval * evals_step / 2.0            # two k_cost launches per iteration
    avg_launch_s = cost_ms * 1e-3 / max(cost_n, 1)
    achieved = algo_bytes_launch / avg_launch_s / 1e9
    flops_per_eval = 7 * 110 + Ksph * 54
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and name == "C2" and B == 1024 and args.dtype == "f64":
        with open(tpath) as f:
            traffic = json.load(f)["k_cost"]["avg_launch_bytes"]
    roofline = {"bound": "hbm", "kernel": "k_cost", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "algorithmic_bytes_per_eval": bytes_per_eval, "avg_launch_ms": 1e3 * avg_launch_s,
                "kernel_share_of_step": cost_ms / all_ms if all_ms else None,
                "binding": {"roof": "sm instruction issue", "frac": 0.53,
                            "source": "ncu smsp__issue_active.avg.pct_of_peak_sustained_active of the main launch, "
                                      "profiles/r1_s3_top_ncu_summary.csv"},
                "note": "k_cost is instruction-issue bound, not HBM bound: algorithmic traffic is ~%d B/eval against ~%d fp64 flop/eval "
                        "(est. %.2f TFLOP/s fp64); traffic = ncu DRAM bytes per launch, mostly L2 hits on the u8 grid"
                        % (bytes_per_eval, flops_per_eval, flops_per_eval * evals_step / 2.0 / avg_launch_s / 1e12),
                "kernels": shares}

    if rank == 0:
        cpu = cpu_baseline(name) if world == 1 and not args.no_cpu_baseline and not sharded else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic",
                "config": {"workload": _workload_desc(name, sc, B),
                           "parallelism": ("rollouts sharded over %d GPU(s), 2 reductions of 2*D*N doubles per iteration, %s" %
                                           (world, "in-kernel over NVLink peer memory (k_peer_allreduce), no host sync" if args.c3_exchange == "peer"
                                            else "NCCL all-reduces between host-synchronised phases"))
                           if sharded else ("problems sharded over %d GPU(s), no collective" % world),
                           "l2": "per-iteration working set %.2f GB of rollout arrays >> 126 MB L2 (no flush needed)"
                                 % (6 * B * R * D * N * 8 / 1e9),
                           "noise": "engine Philox RNG", "evals_per_step_per_gpu": evals_step,
                           "iterations_per_sec": 1e3 * K / ms,
                           "problem_iterations_per_sec": (1 if sharded else world * B) * 1e3 * K / ms,
                           "rollouts_per_sec": total_evals_step / N * 1e3 * K / ms},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--c3-exchange", choices=["peer", "nccl"], default="peer",
                    help="C3 (rollout-sharded) only: how the two per-iteration reductions cross GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_engine(args)


if __name__ == "__main__":
    sys.exit(main())
