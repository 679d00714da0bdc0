"""Pins the CPU oracle to OUTPUTS OF THE REFERENCE ITSELF.

tests/golden/ref_*.npz were produced by 13 of the reference's 14 translation units, compiled UNMODIFIED from
/root/reference/stomp_motion_planner/src (PolicyImprovementLoop, PolicyImprovement, CovariantTrajectoryPolicy,
MultivariateGaussian, StompCost, StompOptimizer incl. execute / handleJointLimits / performForwardKinematics / optimize,
StompTrajectory, StompCollisionPoint, both TreeFkSolverJointPosAxis solvers, OrientationConstraintEvaluator,
StompParameters) against stand-in headers for the third-party packages it does not vendor (oracle/ref_shim/: Eigen 2,
roscpp, Boost, KDL, distance_field, Bullet).  oracle/ref_driver.cpp + tests/golden/make_ref_golden.py are the recipe.

The oracle is fed the noise the reference drew (host-injection mode) and must reproduce every intermediate of
PolicyImprovementLoop::runSingleIteration, every per-sphere quantity of StompOptimizer::execute and the statistics of
StompOptimizer::optimize.

Tolerances: integer work (voxel indices, collision flags, reuse ranking, iteration counts) exact; floating point 1e-8
relative to the array's largest magnitude — the oracle and the compiled reference sum dense products in different orders and
invert R (cond ~ 6e6) with different eliminations; observed agreement is 1e-13 (N=20) to 3e-9 (N=100).
"""
import os

import numpy as np
import pytest

from oracle import oracle, reference
from stomp_motion_planner_icra2011_b200 import _abi
from tests import ref_golden as rg
from tests.helpers import assert_close

RTOL = 1e-8


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_setup_matrices_match_the_compiled_reference(name):
    """R, R^-1, M (column-max scaling), chol(R^-1), the joint-limit Q^-1 and the min-control-cost trajectory."""
    g = rg.load("ref_iter_%s_c1" % name.lower())
    sc = rg.scenario(name, g)
    o = oracle.Oracle(sc, 0)
    for f, nm in rg.SETUP_FIELDS + ((_abi.FIELD_THETA, "theta0"),):
        assert_close(o.get(f), g[nm], 1e-7 if nm == "covariance_cholesky" else RTOL, nm)
    # KAT from src/policy_improvement.cpp:430-436, now checked on the reference's own M
    np.testing.assert_allclose(g["projection_matrix"].max(axis=0), 1.0 / sc.num_time_steps, rtol=1e-12)
    assert float(g["movement_dt"][0]) == sc.movement_duration / (sc.num_time_steps + 1)
    got = o.compute_control_costs(g["cc_parameters"][None], g["cc_noise"][None], 0.5 * sc.smoothness_cost_weight)[0]
    assert_close(got, g["cc_out"], RTOL, "computeControlCosts")


def _check_iterations(sc, g, o):
    R = sc.num_rollouts
    assert_close(o.get(_abi.FIELD_THETA), g["theta0"], RTOL, "theta0")
    for it in rg.iterations(g):
        k = "it%d_" % it
        cost, cf, ngen = o.iterate(it, g[k + "noise"])
        assert ngen == int(g[k + "num_rollouts_gen"])
        np.testing.assert_array_equal(o.get(_abi.FIELD_NOISE)[:ngen], g[k + "noise"])
        for f, nm in rg.ITERATION_FIELDS:
            assert_close(o.get(f), g[k + nm], RTOL, "%s it %d" % (nm, it))
        assert_close(cost, float(g[k + "noiseless_cost"]), RTOL, "noise-less rollout cost")
        assert cf == int(g[k + "noiseless_flags"][0])
        assert_close(o.get(_abi.FIELD_NOISELESS_COSTS), g[k + "exec_costs"][-1], RTOL, "noise-less state costs")
        ecf = o.get(_abi.FIELD_COLLISION_FREE)
        np.testing.assert_array_equal(ecf[:ngen], g[k + "exec_collision_free"][:ngen])
        assert ecf[R] == g[k + "exec_collision_free"][-1]
        # integer work: the ranking that selects the next iteration's reused rollouts
        assert rg.reuse_ranking(o.get(_abi.FIELD_ROLLOUT_TOTAL_COSTS), R) == rg.reuse_ranking(g[k + "totals"], R)
        yield it, k


@pytest.mark.parametrize("name,cumulative", [("tiny", 0), ("tiny", 1), ("C1", 1), ("C1", 0)])
def test_oracle_reproduces_the_reference_iterations(name, cumulative):
    g = rg.load("ref_iter_%s_c%d" % (name.lower(), cumulative))
    sc = rg.scenario(name, g, cumulative)
    assert len(list(_check_iterations(sc, g, oracle.Oracle(sc, 0)))) >= 2


def test_oracle_reproduces_the_reference_iterations_with_orientation_constraints():
    g = rg.load("ref_iter_tiny_constraints")
    sc, cons, w = rg.constraint_scene(g)
    o = oracle.Oracle(sc, 0)
    o.set_constraints(cons, w)
    for it, k in _check_iterations(sc, g, o):
        assert o.get(_abi.FIELD_CONSTRAINTS_SATISFIED)[sc.num_rollouts] == g[k + "noiseless_flags"][1]


def test_reused_rollouts_are_the_references_choice():
    """slots [R_gen, R) of iteration i+1 hold the R_reuse cheapest of iteration i's R rollouts + the noise-less one."""
    g = rg.load("ref_iter_c1_c1")
    R, Rre = 10, 5
    for it in rg.iterations(g)[:-1]:
        order = rg.reuse_ranking(g["it%d_totals" % it], R)[:Rre]
        prev_params = np.concatenate([g["it%d_parameters" % it], g["it%d_theta" % it][None]])
        np.testing.assert_array_equal(g["it%d_parameters" % (it + 1)][R - Rre:], prev_params[order])


def _check_cost_plugin(sc, g, o, constraints=False):
    org, res = sc.sdf.origin, sc.sdf.resolution
    n = g["parameters"].shape[0]
    flips = 0
    for r in range(n):
        for i, itn in enumerate((1, 2)):
            costs, cf = o.execute(g["parameters"][r], itn)
            assert_close(costs[0], g["costs"][r, i], RTOL, "state costs, rollout %d iteration %d" % (r, itn))
            assert cf[0] == g["flags"][r, i, 0]
            if constraints:
                assert o.execute_constraints_satisfied(1)[0] == g["flags"][r, i, 1]
        dbg, clipped = o.execute_debug(g["parameters"][r])
        assert_close(clipped, g["clipped"][r], RTOL, "trajectory after handleJointLimits")
        assert_close(dbg["position"], g["dbg_position"][r], 1e-9, "collision point positions", atol_scale=1e-9)
        safe = rg.boundary_safe(g["dbg_position"][r], org, res)
        flips += int((~safe).sum())
        np.testing.assert_array_equal(dbg["voxel"][safe], g["dbg_voxel"][r][safe])          # bit-exact integer work
        np.testing.assert_array_equal(dbg["in_collision"][safe], g["dbg_in_collision"][r][safe])
        assert_close(dbg["potential"][safe], g["dbg_potential"][r][safe], RTOL, "potential")
        assert_close(dbg["vel_mag"], g["dbg_vel_mag"][r], RTOL, "velocity magnitude")
    assert flips < 5


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_cost_plugin_matches_the_compiled_stomp_optimizer(name):
    """StompOptimizer::execute: joint limits, FK, sphere positions, voxel indices, collision flags, potential, |v|, costs."""
    g = rg.load("ref_cost_%s" % name.lower())
    sc = rg.scenario(name, g)
    _check_cost_plugin(sc, g, oracle.Oracle(sc, 0))
    assert g["dbg_in_collision"].any() and (g["dbg_potential"] > 0).any(), "fixture exercises no obstacle"
    assert np.abs(g["clipped"] - g["parameters"]).max() > 1e-3, "fixture exercises no joint limit"


def test_torque_term_matches_the_compiled_get_torques():
    """StompOptimizer::getTorques + the sum of |tau_j| (the RNE solver itself is third-party on both sides: tests/test_torque_cpu.py)."""
    from tests.golden.make_ref_golden import TORQUE_WEIGHT
    g, plain = rg.load("ref_cost_tiny_torque"), None
    sc = rg.scenario("tiny", g)
    o = oracle.Oracle(sc, 0)
    o.set_dynamics(TORQUE_WEIGHT)
    _check_cost_plugin(sc, g, o)
    plain = oracle.Oracle(sc, 0).execute(g["parameters"][0], 2)[0][0]
    assert (g["costs"][0, 1] - plain).min() > 1e-3, "fixture's torque term is not active"
    gi = rg.load("ref_iter_tiny_torque")
    o2 = oracle.Oracle(rg.scenario("tiny", gi), 0)
    o2.set_dynamics(TORQUE_WEIGHT)
    assert len(list(_check_iterations(sc, gi, o2))) == 3


def test_cost_plugin_with_orientation_constraints_matches_the_compiled_reference():
    g = rg.load("ref_cost_tiny_constraints")
    sc, cons, w = rg.constraint_scene(g)
    o = oracle.Oracle(sc, 0)
    o.set_constraints(cons, w)
    _check_cost_plugin(sc, g, o, constraints=True)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_forward_kinematics_of_random_trees_matches_the_compiled_solvers(seed):
    """branching trees with fixed / prismatic / revolute joints through TreeFkSolverJointPosAxisPartial."""
    g = rg.load("ref_cost_tree%d" % seed)
    sc = rg.tree_scene(seed, g)
    _check_cost_plugin(sc, g, oracle.Oracle(sc, 0))


@pytest.mark.parametrize("stem,name", [("ref_optimize_tiny_s8", "tiny"), ("ref_optimize_tiny_s7", "tiny"), ("ref_optimize_c1_s8", "C1")])
def test_optimize_bookkeeping_matches_the_compiled_reference(stem, name):
    """StompOptimizer::optimize end to end: success iteration, early exit, cost log, best trajectory."""
    g = rg.load(stem)
    sc = rg.scenario(name, g, seed=int(g["seed"]))
    o = oracle.Oracle(sc, 0)
    assert_close(o.get(_abi.FIELD_THETA), g["theta0"], RTOL, "theta0")

    def step(it):
        cost, cf, _ = o.iterate(it, g["it%d_noise" % it])
        return cost, cf, 1, o.get_parameters()

    res = rg.host_bookkeeping(step, int(g["max_iterations"]), int(g["max_iterations_after_collision_free"]))
    assert res["success"] == bool(g["stats_success"])
    assert res["success_iteration"] == int(g["stats_success_iteration"]) == int(g["stats_collision_success_iteration"])
    assert res["iterations"] == int(g["stats_iterations"])
    assert res["last_improvement_iteration"] == int(g["stats_last_improvement_iteration"])
    assert_close(res["costs"], g["stats_costs"], 1e-7, "STOMPStatistics.costs")
    assert_close(res["best_cost"], float(g["stats_best_cost"]), 1e-7, "best cost")


@pytest.mark.skipif(not os.path.isdir(reference.REFERENCE_ROOT), reason="/root/reference is not on this machine")
def test_fixtures_are_what_the_compiled_reference_produces_today():
    """Rebuild-and-rerun check (only where the reference sources exist): the committed vectors are reproducible."""
    from tests.golden import make_ref_golden
    for stem, fresh in (("ref_iter_tiny_c1", make_ref_golden.iteration_case("tiny", 1, 5)),
                        ("ref_cost_tiny", make_ref_golden.cost_plugin_case("tiny", 4, 21))):
        g = rg.load(stem)
        assert sorted(fresh) == sorted(g.files)
        for k in g.files:
            np.testing.assert_allclose(fresh[k], g[k], rtol=1e-12, atol=1e-300, err_msg=k)


def test_collision_point_generation_matches_the_compiled_robot_model():
    """StompRobotModel::generateLinkCollisionPoints / generateAttachedObjectCollisionPoints / populatePlanningGroupCollisionPoints
    (src/stomp_robot_model.cpp:265-496) against Robot.add_link_spheres / add_attached_object, which urdf.py builds on."""
    from stomp_motion_planner_icra2011_b200 import scenes
    from tests.golden.make_ref_golden import ATTACHED
    g = rg.load("ref_collision_points")
    as_array = lambda rb: np.array([[s["segment"], s["radius"], s["clearance"], *s["pos"]] for s in rb.spheres])  # noqa: E731
    # the synthetic arm's own table is what the reference generates from its tree: bit-identical
    np.testing.assert_array_equal(as_array(scenes.pr2_right_arm()), g["plain_points"])
    # per-link clearances, a link outside the group's reach (dropped), attached objects with padding
    rb = scenes.pr2_right_arm()
    rb.spheres = []
    names = [s["name"] for s in rb.segments]
    for seg, radius, clearance, extension in g["links"]:
        # link points first, then the link's attached-object points, link by link (populatePlanningGroupCollisionPoints)
        rb.add_link_spheres(int(seg), radius, 0.07 if clearance < 0 else clearance, extension)
        for n, shape, dims, pos in ATTACHED:
            if names.index(n) == int(seg):
                rb.add_attached_object(int(seg), shape, dims, pos, padding=0.01, clearance=0.07)
    rb.spheres = [s for s in rb.spheres if rb.moved_by_group(s["segment"])]
    got = as_array(rb)
    assert got.shape == g["points"].shape
    np.testing.assert_array_equal(got[:, 0], g["points"][:, 0])
    np.testing.assert_allclose(got, g["points"], rtol=1e-14, atol=0)
    assert len(g["points"]) > len(g["plain_points"]) and not (g["points"][:, 0] == names.index("torso_lift_link")).any()


def test_collision_object_rasterisation_matches_the_compiled_collision_space():
    """StompCollisionSpace::addCollisionObjectsToPoints (src/stomp_collision_space.cpp:198-297), lattice loops with accumulated
    `x += resolution_`, rotated boxes and cylinders: the occupied cells are integer work, bit-exact.  (The engine's
    stomp_engine_build_sdf is held to oracle/sdf_builder.py on the same scene by tests/test_gpu_parity.py and to this fixture by
    tests/test_gpu_reference_golden.py.)"""
    from oracle import sdf_builder
    from tests.golden.make_ref_golden import SDF_SCENE, map_points, sdf_objects
    boxes, cyls = sdf_objects()
    for stem, pts in (("ref_collision_cells", None), ("ref_collision_cells_points", map_points())):   # + a collision map
        g = rg.load(stem)
        _, occ = sdf_builder.build(boxes=boxes, cylinders=cyls, max_distance=0.17, points=pts, **SDF_SCENE)
        want = np.unpackbits(g["occupancy_bits"])[:occ.size].reshape(g["shape"]).astype(bool)
        assert occ.shape == want.shape and int(want.sum()) == int(g["num_occupied"]) > 10000
        np.testing.assert_array_equal(occ, want)
    assert int(rg.load("ref_collision_cells_points")["num_occupied"]) > int(rg.load("ref_collision_cells")["num_occupied"]) + 500
