"""Generates tests/golden/ref_pi2_*.npz by RUNNING THE REFERENCE'S OWN CODE: the unmodified translation units
src/policy_improvement.cpp, src/policy_improvement_loop.cpp, src/covariant_trajectory_policy.cpp and src/stomp_cost.cpp
of /root/reference/stomp_motion_planner, compiled against the stand-in headers of oracle/ref_shim/ into
oracle/_ref/libstomp_ref_pi2.so (oracle/Makefile target `ref`, oracle/ref_driver.cpp).

Run from the repo root in a container that has /root/reference:  python tests/golden/make_ref_golden.py

Everything in the files is an output of the reference's PolicyImprovementLoop::runSingleIteration — the noise its
MultivariateGaussian drew, M eps, control costs, cumulative costs, probabilities, updates, the updated policy, getCost()
of every rollout (which fixes the reuse order) — except `state_costs`: the reference's cost plugin
(StompOptimizer::execute) cannot be compiled here, so Task::execute is served by the CPU restatement's cost plugin
(oracle/stomp_oracle.cpp) and its answers are recorded next to the parameters they were asked for (`exec_*`).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_pi2 as rp  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from stomp_motion_planner_icra2011_b200 import scenes  # noqa: E402

CASES = (("tiny", 0, 5), ("tiny", 1, 5), ("C1", 1, 4), ("C1", 0, 2))
PER_ROLLOUT = ("noise", "parameters", "noise_projected", "control_costs", "cumulative_costs", "probabilities")


def run_case(name, cumulative, iterations):
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=cumulative)
    D, N, R = sc.robot.num_dimensions, sc.num_time_steps, sc.num_rollouts
    plugin = Oracle(sc, 0)
    ref = rp.ReferencePI2(N, D, R, sc.num_reused_rollouts, sc.movement_duration, sc.ridge_factor, sc.derivative_costs,
                          sc.noise_stddev, sc.noise_decay, sc.smoothness_cost_weight, cumulative, sc.start[0], sc.goal[0],
                          lambda p, it: plugin.execute(p, it)[0][0])
    out = {"theta0": ref.get_parameters(), "parameters_all0": ref.get("parameters_all"),
           "movement_dt": ref.get("movement_dt")}
    if cumulative == 1:  # setup matrices do not depend on the flag; store them once per scene
        for f in ("control_cost_matrix", "inv_control_cost_matrix", "projection_matrix", "covariance_cholesky"):
            out[f] = ref.get(f)
        out["quad_cost_inv"] = rp.quad_cost_inv(N + 12, sc.discretization, sc.derivative_costs, sc.ridge_factor, np.ones(D))[0]
        rng = np.random.default_rng(5)
        p, e = rng.standard_normal((D, N)), 0.1 * rng.standard_normal((D, N))
        out["cc_parameters"], out["cc_noise"] = p, e
        out["cc_out"] = ref.compute_control_costs(p, e, 0.5 * sc.smoothness_cost_weight)
    for it in range(1, iterations + 1):
        ref.calls.clear()
        ref.run_single_iteration(it)
        k = "it%d_" % it
        ngen = int(ref.get("num_rollouts_gen")[0])
        out[k + "num_rollouts_gen"] = np.array(ngen, dtype=np.int32)
        for f in PER_ROLLOUT:
            a = ref.get(f)
            out[k + f] = a[:ngen] if f == "noise" else a
        out[k + "state_costs"] = ref.get("state_costs")
        out[k + "totals"] = np.concatenate([ref.get("total"), ref.get("extra_total")])
        out[k + "extra_control_costs"] = ref.get("extra_control_costs")[0]
        out[k + "updates"] = ref.get("parameter_updates")
        out[k + "theta"] = ref.get_parameters()
        assert len(ref.calls) == ngen + 1 and all(c[0] == it for c in ref.calls)
        out[k + "exec_costs"] = np.stack([c[2] for c in ref.calls])  # answers of the cost plugin, last = noise-less rollout
    ref.close()
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, cumulative, iterations in CASES:
        out = run_case(name, cumulative, iterations)
        path = os.path.join(here, "ref_pi2_%s_c%d.npz" % (name.lower(), cumulative))
        np.savez_compressed(path, **out)
        print(path, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
