"""Shared by tests/test_reference_pinning.py (CPU oracle) and tests/test_gpu_reference_golden.py (CUDA engine): loading of
the vectors produced by the reference's own compiled code (tests/golden/make_ref_golden.py) and the scenarios they were
produced on."""
import os

import numpy as np

from stomp_motion_planner_icra2011_b200 import _abi, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ITERATION_FIELDS = ((_abi.FIELD_NOISE_PROJECTED, "noise_projected"), (_abi.FIELD_PARAMETERS, "parameters"),
                    (_abi.FIELD_STATE_COSTS, "state_costs"), (_abi.FIELD_CONTROL_COSTS, "control_costs"),
                    (_abi.FIELD_CUMULATIVE_COSTS, "cumulative_costs"), (_abi.FIELD_PROBABILITIES, "probabilities"),
                    (_abi.FIELD_UPDATES, "updates"), (_abi.FIELD_THETA, "theta"), (_abi.FIELD_ROLLOUT_TOTAL_COSTS, "totals"))
SETUP_FIELDS = ((_abi.FIELD_CONTROL_COST, "control_cost_matrix"), (_abi.FIELD_INV_CONTROL_COST, "inv_control_cost_matrix"),
                (_abi.FIELD_PROJECTION, "projection_matrix"), (_abi.FIELD_NOISE_CHOLESKY, "covariance_cholesky"),
                (_abi.FIELD_QUAD_COST_INV, "quad_cost_inv"))


def load(stem):
    return np.load(os.path.join(GOLDEN, stem + ".npz"))


def iterations(g):
    return sorted(int(k[2:k.index("_")]) for k in g.files if k.startswith("it") and k.endswith("_noise"))


def scenario(name, g, cumulative=1, **kw):
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=cumulative, **kw)
    sc.movement_duration = float(g["movement_duration"])   # what the reference derived from the trajectory it was given
    return sc


def constraint_scene(g):
    from tests.golden.make_ref_golden import constraint_scene as cs
    sc, cons, w = cs()
    sc.movement_duration = float(g["movement_duration"])
    return sc, cons, w


def tree_scene(seed, g):
    rng = np.random.default_rng(100 + seed)
    sc = scenario("tiny", g)
    sc.robot = scenes.random_tree(rng)
    sc.start, sc.goal = g["start"], g["goal"]
    return sc


def reuse_ranking(totals, R):
    """the (getCost(), index) order of src/policy_improvement.cpp:178-196; the extra rollout sorts as index -1."""
    return sorted(range(R + 1), key=lambda r: (totals[r], -1 if r == R else r))


def boundary_safe(position, origin, resolution, margin=1e-6):
    """spheres whose reference position is not within `margin` cells of a rounding boundary of int(round(.)): only there may a
    correct implementation whose positions differ by 1e-10 m land in the neighbouring voxel."""
    frac = (position - np.asarray(origin)) / resolution
    return np.all(np.abs(np.abs(frac - np.round(frac)) - 0.5) > margin, axis=-1)


def host_bookkeeping(step, max_iterations, max_cf):
    """StompOptimizer::optimize's bookkeeping (src/stomp_optimizer.cpp:284-359) around step(it) -> (cost, collision_free,
    constraints_satisfied, trajectory)."""
    cf_count, succ, last_imp, best_cost, best, costs = 0, -1, -1, 0.0, None, []
    it = 0
    while it < max_iterations:
        cost, cf, cs, traj = step(it + 1)
        cf_count = cf_count + 1 if (cf and cs) else 0
        if cf and cs and succ == -1:
            succ = it
        costs.append(cost)
        if it == 0:
            best_cost, best = cost, traj
        elif cost < best_cost and cf and cs:
            best_cost, best, last_imp = cost, traj, it
        if cf_count >= max_cf:
            it += 1
            break
        it += 1
    return dict(success=succ >= 0, success_iteration=succ, iterations=len(costs), last_improvement_iteration=last_imp,
                best_cost=best_cost, best_trajectory=best, costs=np.array(costs))
