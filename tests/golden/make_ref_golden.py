"""Generates tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN CODE: 13 of its 14 translation units
(policy_improvement, policy_improvement_loop, covariant_trajectory_policy, stomp_cost, stomp_optimizer, stomp_trajectory,
stomp_collision_point, treefksolverjointposaxis, treefksolverjointposaxis_partial, constraint_evaluator,
stomp_parameters, stomp_robot_model, stomp_collision_space; the 14th is the ROS node's main), compiled unmodified from /root/reference/stomp_motion_planner/src against the stand-in headers of
oracle/ref_shim/ into oracle/_ref/libstomp_ref.so (oracle/Makefile target `ref`, oracle/ref_driver.cpp).

Run from the repo root in a container that has /root/reference:  python tests/golden/make_ref_golden.py

Every array is an output of the reference: the noise its MultivariateGaussian drew, the rollouts it executed, the state
costs / collision flags / sphere positions / potentials of StompOptimizer::execute, M eps, control and cumulative costs,
probabilities, updates, the updated policy, getCost() of every rollout, and the statistics StompOptimizer::optimize
publishes.  Inputs are the synthetic scenes of stomp_motion_planner_icra2011_b200/scenes.py (same seeds as the tests).

What the stand-in headers restate instead of the reference (third-party packages the reference does not vendor): Eigen 2
dense primitives, KDL frame algebra, the distance_field cell lookup, Bullet's getRPY (see the headers of oracle/ref_shim/).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference as rp  # noqa: E402
from stomp_motion_planner_icra2011_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ITERATION_CASES = (("tiny", 0, 5), ("tiny", 1, 5), ("C1", 1, 4), ("C1", 0, 2))
PER_ROLLOUT = ("noise", "parameters", "noise_projected", "control_costs", "cumulative_costs", "probabilities")


def scenario(name, cumulative=1, **kw):
    """the test scenario, with the policy duration the reference itself derives (int-truncated group-trajectory duration,
    src/stomp_optimizer.cpp:185 / include/stomp_motion_planner/stomp_trajectory.h:270-273)."""
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=cumulative, **kw)
    sc.movement_duration = float(int((sc.num_time_steps + 11) * sc.discretization))
    return sc


def constraint_scene():
    sc = scenario("tiny")
    seg = [i for i, g in enumerate(sc.robot.segments) if g["name"] == "r_gripper_palm_link"][0]
    cons = [dict(segment=seg, body_fixed=0, orientation=(0.0, 0.0, 0.0, 1.0), tolerances=(0.4, 0.3, 4.0), weight=1.5),
            dict(segment=seg - 2, body_fixed=1, orientation=(0.1, -0.2, 0.3, 0.9), tolerances=(4.0, 0.5, 0.6), weight=0.7)]
    return sc, cons, 2.0


TORQUE_WEIGHT = 0.02     # the shipped configurations use 0 (term off); this switches StompOptimizer::getTorques on


def iteration_case(name, cumulative, iterations, constraints=(), weight=0.0, torque=0.0):
    sc = scenario(name, cumulative)
    D, N, R = sc.robot.num_dimensions, sc.num_time_steps, sc.num_rollouts
    ref = rp.ReferenceOptimizer(sc, 0, constraints=constraints, constraint_cost_weight=weight)
    assert ref.movement_duration == sc.movement_duration
    if torque:
        ref.set_dynamics(torque)
    ref.begin()
    out = {"movement_duration": np.array(ref.movement_duration), "theta0": ref.get("theta"),
           "parameters_all0": ref.get("parameters_all"), "movement_dt": ref.get("movement_dt")}
    if cumulative == 1:  # setup matrices do not depend on the flag; store them once per scene
        for f in ("control_cost_matrix", "inv_control_cost_matrix", "projection_matrix", "covariance_cholesky", "quad_cost_inv"):
            out[f] = ref.get(f)
        # Policy::computeControlCosts on arbitrary inputs, through the PI^2-only entry points of the same library
        pi = rp.ReferencePI2(N, D, R, sc.num_reused_rollouts, sc.movement_duration, sc.ridge_factor, sc.derivative_costs,
                             sc.noise_stddev, sc.noise_decay, sc.smoothness_cost_weight, cumulative, sc.start[0], sc.goal[0],
                             lambda p, it: np.zeros(N))
        rng = np.random.default_rng(5)
        out["cc_parameters"], out["cc_noise"] = rng.standard_normal((D, N)), 0.1 * rng.standard_normal((D, N))
        out["cc_out"] = pi.compute_control_costs(out["cc_parameters"], out["cc_noise"], 0.5 * sc.smoothness_cost_weight)
        pi.close()
    for it in range(1, iterations + 1):
        ref.get("exec_clear")
        cost, cf, cs = ref.iterate(it)
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
        out[k + "theta"] = ref.get("theta")
        out[k + "noiseless_cost"] = np.array(cost)
        out[k + "noiseless_flags"] = np.array([cf, cs], dtype=np.int32)
        assert int(ref.get("exec_count")[0]) == ngen + 1
        out[k + "exec_costs"] = ref.get("exec_costs")            # answers of StompOptimizer::execute, last = noise-less rollout
        out[k + "exec_collision_free"] = ref.get("exec_collision_free").astype(np.int32)
    ref.close()
    return out


def cost_plugin_case(name, rollouts, seed, constraints=(), weight=0.0, robot=None, start=None, goal=None, torque=0.0):
    """StompOptimizer::execute on noisy rollouts around the min-control-cost trajectory + its per-sphere internals."""
    sc = scenario(name)
    if robot is not None:
        sc.robot, sc.start, sc.goal = robot, start, goal
    D, N = sc.robot.num_dimensions, sc.num_time_steps
    ref = rp.ReferenceOptimizer(sc, 0, constraints=constraints, constraint_cost_weight=weight)
    if torque:
        ref.set_dynamics(torque)
    ref.begin()
    L, theta = ref.get("covariance_cholesky"), ref.get("theta")
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((rollouts, D, N))
    scale = 2.0 * (1.0 + 3.0 * np.arange(rollouts))             # later rollouts leave the joint limits
    params = theta[None] + np.einsum("ij,rdj->rdi", L, z) * scale[:, None, None]
    out = {"movement_duration": np.array(ref.movement_duration), "parameters": params}
    costs, flags, dbg, clipped = [], [], [], []
    for r in range(rollouts):
        for it in (1, 2):
            c, cf, cs = ref.execute(params[r], it)
            costs.append(c)
            flags.append((cf, cs))
        d, cl = ref.execute_debug(params[r])
        dbg.append(d)
        clipped.append(cl)
    out["costs"] = np.stack(costs).reshape(rollouts, 2, N)        # [r][iteration 1 | 2][N]
    out["flags"] = np.array(flags, dtype=np.int32).reshape(rollouts, 2, 2)   # collision_free, constraints_satisfied
    for key in ("voxel", "in_collision", "position", "potential", "vel_mag"):
        out["dbg_" + key] = np.stack([d[key] for d in dbg])
    out["clipped"] = np.stack(clipped)
    ref.close()
    return out


def optimize_case(name, seed, max_iterations, max_cf):
    sc = scenario(name, seed=seed)
    D, N, R = sc.robot.num_dimensions, sc.num_time_steps, sc.num_rollouts
    ref = rp.ReferenceOptimizer(sc, 0, max_iterations=max_iterations, max_iterations_after_collision_free=max_cf)
    theta0 = ref.get("theta")
    res = ref.optimize()
    params = ref.get("exec_parameters")
    n_it = res["iterations"]
    out = {"movement_duration": np.array(ref.movement_duration), "seed": np.array(seed), "max_iterations": np.array(max_iterations),
           "max_iterations_after_collision_free": np.array(max_cf), "theta0": theta0}
    for k in ("success", "success_iteration", "collision_success_iteration", "best_cost", "iterations",
              "last_improvement_iteration", "costs", "best_trajectory"):
        out["stats_" + k] = np.asarray(res[k])
    # the rollouts executed per iteration: R new ones in the first iteration, R - R_reuse afterwards, then the noise-less one
    pos, theta = 0, theta0
    for it in range(1, n_it + 1):
        ngen = R if it == 1 else R - sc.num_reused_rollouts
        out["it%d_noise" % it] = params[pos:pos + ngen] - theta[None]     # Rollout::noise_ up to one rounding of theta + eps
        theta = params[pos + ngen]
        pos += ngen + 1
    assert pos == len(params)
    out["exec_collision_free"] = ref.get("exec_collision_free").astype(np.int32)
    ref.close()
    return out


def random_tree_case(seed):
    rng = np.random.default_rng(100 + seed)
    rb = scenes.random_tree(rng)
    start = rng.uniform(-1, 1, (1, rb.num_dimensions))
    goal = rng.uniform(-1, 1, (1, rb.num_dimensions))
    out = cost_plugin_case("tiny", 1, 40 + seed, robot=rb, start=start, goal=goal)
    out["start"], out["goal"] = start, goal
    return out


PR2_LINKS = (("r_upper_arm_link", 0.10, 0.0), ("r_forearm_link", 0.065, 0.0), ("r_gripper_palm_link", 0.06, 0.0),
             ("r_gripper_l_finger_link", 0.03, 0.01), ("r_gripper_l_finger_tip_link", 0.03, 0.01),
             ("r_gripper_r_finger_link", 0.03, 0.01), ("r_gripper_r_finger_tip_link", 0.03, 0.01))   # link, link_radius, link_extension
ATTACHED = (("r_gripper_palm_link", "box", (0.10, 0.06, 0.20), (0.12, 0.0, 0.0)),
            ("r_gripper_palm_link", "cylinder", (0.03, 0.25), (0.10, 0.02, -0.01)),
            ("r_forearm_link", "sphere", (0.05,), (0.2, 0.0, 0.04)))
SDF_SCENE = dict(size=(2.0, 3.0, 2.2), origin=(-0.5, -1.5, -0.3), resolution=0.015)


def sdf_objects():
    ident = (0.0, 0.0, 0.0, 1.0)
    yaw = (0.0, 0.0, float(np.sin(0.35)), float(np.cos(0.35)))
    tilt = (float(np.sin(0.2)) * 0.6, float(np.sin(0.2)) * 0.8, 0.0, float(np.cos(0.2)))
    boxes = [((0.8, -0.1, 0.015), ident, (0.4, 1.2, 0.03)), ((0.8, -0.685, 0.8), ident, (0.4, 0.03, 1.6)),
             ((0.3, 0.6, 0.9), yaw, (0.25, 0.1, 0.3)), ((-0.45, -1.45, 1.0), tilt, (0.3, 0.3, 0.3))]   # last one pokes out of the grid
    cyls = [((0.62, -0.62, 0.6), ident, 0.1, 1.2), ((0.2, 0.2, 0.4), tilt, 0.07, 0.5)]
    return boxes, cyls


def collision_point_case():
    """StompRobotModel::generateLinkCollisionPoints / generateAttachedObjectCollisionPoints / populatePlanningGroupCollisionPoints."""
    rb = scenes.pr2_right_arm()
    rb.spheres = []
    names = [g["name"] for g in rb.segments]
    links = [(names.index(n), r, None if i % 2 else 0.05 + 0.01 * i, e) for i, (n, r, e) in enumerate(PR2_LINKS)]
    # a link no group joint moves: the planning group drops its points (src/stomp_robot_model.cpp:308-334)
    links.append((names.index("torso_lift_link"), 0.2, None, 0.0))
    attached = [(names.index(n), shape, dims, pos) for n, shape, dims, pos in ATTACHED]
    pts = rp.collision_points(rb, links, default_clearance=0.07, attached=attached, attached_padding=0.01)
    as_array = lambda p: np.array([[a[0], a[1], a[2], *a[3]] for a in p])  # noqa: E731
    out = {"links": np.array([[l[0], l[1], -1.0 if l[2] is None else l[2], l[3]] for l in links]), "points": as_array(pts)}
    # the plain configuration of the synthetic arm (every clearance the default, no attached objects)
    plain = [(names.index(n), r, None, e) for n, r, e in PR2_LINKS]
    out["plain_points"] = as_array(rp.collision_points(rb, plain, default_clearance=0.07))
    return out


def map_points():
    """a synthetic collision map: points on a tilted plane patch, some outside the grid"""
    rng = np.random.default_rng(17)
    uv = rng.uniform(-0.4, 0.4, (4000, 2))
    return np.stack([0.9 + uv[:, 0], 1.2 + uv[:, 1] * 1.2, 1.0 + 0.3 * uv[:, 0] - 0.2 * uv[:, 1]], axis=1)


def collision_cells_case(with_points=False):
    """StompCollisionSpace::addCollisionObjectsToPoints + the distance field's cell binning."""
    boxes, cyls = sdf_objects()
    occ, npts = rp.collision_object_cells(boxes=boxes, cylinders=cyls, points=map_points() if with_points else None, **SDF_SCENE)
    return {"occupancy_bits": np.packbits(occ.ravel()), "shape": np.array(occ.shape), "num_points": np.array(npts),
            "num_occupied": np.array(int(occ.sum()))}


def main():
    def save(stem, out):
        path = os.path.join(HERE, stem + ".npz")
        np.savez_compressed(path, **out)
        print(path, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")

    for name, cumulative, iterations in ITERATION_CASES:
        save("ref_iter_%s_c%d" % (name.lower(), cumulative), iteration_case(name, cumulative, iterations))
    sc, cons, w = constraint_scene()
    save("ref_iter_tiny_constraints", iteration_case("tiny", 1, 3, cons, w))
    save("ref_cost_tiny", cost_plugin_case("tiny", 4, 21))
    save("ref_cost_c1", cost_plugin_case("C1", 2, 22))
    save("ref_cost_tiny_constraints", cost_plugin_case("tiny", 3, 23, cons, w))
    for seed in (0, 1, 2):
        save("ref_cost_tree%d" % seed, random_tree_case(seed))
    save("ref_cost_tiny_torque", cost_plugin_case("tiny", 3, 24, torque=TORQUE_WEIGHT))
    save("ref_iter_tiny_torque", iteration_case("tiny", 1, 3, torque=TORQUE_WEIGHT))
    save("ref_collision_points", collision_point_case())
    save("ref_collision_cells", collision_cells_case())
    save("ref_collision_cells_points", collision_cells_case(with_points=True))
    save("ref_optimize_tiny_s8", optimize_case("tiny", 8, 60, 5))     # collision free at iteration 28, early exit after 33
    save("ref_optimize_tiny_s7", optimize_case("tiny", 7, 20, 6))     # never collision free: runs to max_iterations
    save("ref_optimize_c1_s8", optimize_case("C1", 8, 40, 5))        # collision free from the first iteration


if __name__ == "__main__":
    main()
