"""GPU parity against OUTPUTS OF THE REFERENCE ITSELF (tests/golden/ref_*.npz: 13 of the reference's 14 translation units
compiled unmodified and run on the synthetic scenes — see tests/test_reference_pinning.py and
tests/golden/make_ref_golden.py).  No oracle is involved at run time.

The engine is driven through the C ABI with the noise the reference's MultivariateGaussian drew (host-injection mode,
BASELINE north_star) and must land on the reference's intermediates, state costs, collision flags, voxel indices, updated
policy and optimize() statistics: integer work exact, floating point within 1e-5 relative (fp64 mode).
"""
import numpy as np
import pytest

from stomp_motion_planner_icra2011_b200 import _abi
from tests import ref_golden as rg
from tests.helpers import RTOL_F64, assert_close, band_solves_only

pytestmark = pytest.mark.gpu


def _engine(sc, band_only=False, **kw):
    from stomp_motion_planner_icra2011_b200.engine import Engine
    if band_only:
        with band_solves_only():
            return Engine(sc, **kw)
    return Engine(sc, **kw)


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_engine_setup_matches_the_compiled_reference(name):
    g = rg.load("ref_iter_%s_c1" % name.lower())
    sc = rg.scenario(name, g)
    eng = _engine(sc)
    for f, nm in rg.SETUP_FIELDS:
        assert_close(np.asarray(eng.get(f)).reshape(g[nm].shape), g[nm], RTOL_F64, nm)
    assert_close(eng.get(_abi.FIELD_THETA)[0], g["theta0"], RTOL_F64, "min-control-cost trajectory")
    got = eng.compute_control_costs(g["cc_parameters"][None, None], g["cc_noise"][None, None], 0.5 * sc.smoothness_cost_weight)
    assert_close(np.asarray(got).reshape(g["cc_out"].shape), g["cc_out"], RTOL_F64, "computeControlCosts")


def _check_iterations(sc, g, eng):
    R = sc.num_rollouts
    for it in rg.iterations(g):
        k = "it%d_" % it
        eng.inject_noise(g[k + "noise"][None])
        cost, cf, ngen = eng.iterate(it)
        assert ngen == int(g[k + "num_rollouts_gen"])
        for f, nm in rg.ITERATION_FIELDS:
            assert_close(eng.get(f)[0], g[k + nm], RTOL_F64, "%s it %d" % (nm, it))
        assert_close(cost[0], float(g[k + "noiseless_cost"]), RTOL_F64, "noise-less rollout cost")
        assert cf[0] == int(g[k + "noiseless_flags"][0])
        assert_close(eng.get(_abi.FIELD_NOISELESS_COSTS)[0], g[k + "exec_costs"][-1], RTOL_F64, "noise-less state costs")
        ecf = eng.get(_abi.FIELD_COLLISION_FREE)[0]
        np.testing.assert_array_equal(ecf[:ngen], g[k + "exec_collision_free"][:ngen])
        assert ecf[R] == g[k + "exec_collision_free"][-1]
        # integer work: the (getCost(), index) ranking that selects next iteration's reused rollouts
        assert rg.reuse_ranking(eng.get(_abi.FIELD_ROLLOUT_TOTAL_COSTS)[0], R) == rg.reuse_ranking(g[k + "totals"], R)
        yield it, k


@pytest.mark.parametrize("band_only", [False, True], ids=["dense", "band"])   # both generation kernels (see helpers.band_solves_only)
@pytest.mark.parametrize("name,cumulative", [("tiny", 0), ("tiny", 1), ("C1", 1), ("C1", 0)])
def test_engine_reproduces_the_reference_iterations(name, cumulative, band_only):
    g = rg.load("ref_iter_%s_c%d" % (name.lower(), cumulative))
    sc = rg.scenario(name, g, cumulative)
    assert len(list(_check_iterations(sc, g, _engine(sc, band_only=band_only, keep_intermediates=1)))) >= 2


def test_engine_reproduces_the_reference_iterations_with_orientation_constraints():
    g = rg.load("ref_iter_tiny_constraints")
    sc, cons, w = rg.constraint_scene(g)
    eng = _engine(sc, keep_intermediates=1)
    eng.set_constraints(cons, w)
    for it, k in _check_iterations(sc, g, eng):
        assert eng.get(_abi.FIELD_CONSTRAINTS_SATISFIED)[0][sc.num_rollouts] == g[k + "noiseless_flags"][1]


def _check_cost_plugin(sc, g, eng, constraints=False):
    org, res = np.asarray(sc.sdf.origin), sc.sdf.resolution
    params = g["parameters"]
    n = params.shape[0]
    for i, itn in enumerate((1, 2)):
        costs, cf = eng.execute(params[None], itn)
        assert_close(costs[0], g["costs"][:, i], RTOL_F64, "state costs, iteration %d" % itn)
        np.testing.assert_array_equal(cf[0], g["flags"][:, i, 0])
        if constraints:
            np.testing.assert_array_equal(eng.execute_constraints_satisfied(n)[0], g["flags"][:, i, 1])
    flips = 0
    for r in range(n):
        dbg = eng.execute_debug(params[r])
        # voxel indices are exactly int(round((x - origin) / res)) of the engine's own positions ...
        np.testing.assert_array_equal(dbg["voxel"], np.round((dbg["position"] - org) / res).astype(np.int32))
        # ... and the reference's wherever its position is not within 1e-6 cells of a rounding boundary
        safe = rg.boundary_safe(g["dbg_position"][r], org, res)
        flips += int((~safe).sum())
        np.testing.assert_array_equal(dbg["voxel"][safe], g["dbg_voxel"][r][safe])
        np.testing.assert_array_equal(dbg["in_collision"][safe], g["dbg_in_collision"][r][safe])
        assert_close(dbg["position"], g["dbg_position"][r], 1e-8, "collision point positions", atol_scale=1e-8)
        assert_close(dbg["potential"][safe], g["dbg_potential"][r][safe], RTOL_F64, "potential")
        assert_close(dbg["vel_mag"], g["dbg_vel_mag"][r], RTOL_F64, "velocity magnitude")
    assert flips < 5


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_engine_cost_plugin_matches_the_compiled_stomp_optimizer(name):
    g = rg.load("ref_cost_%s" % name.lower())
    sc = rg.scenario(name, g)
    _check_cost_plugin(sc, g, _engine(sc))


def test_engine_cost_plugin_with_orientation_constraints_matches_the_compiled_reference():
    g = rg.load("ref_cost_tiny_constraints")
    sc, cons, w = rg.constraint_scene(g)
    eng = _engine(sc)
    eng.set_constraints(cons, w)
    _check_cost_plugin(sc, g, eng, constraints=True)


def test_engine_torque_term_matches_the_compiled_get_torques():
    from tests.golden.make_ref_golden import TORQUE_WEIGHT
    g = rg.load("ref_cost_tiny_torque")
    sc = rg.scenario("tiny", g)
    eng = _engine(sc)
    eng.set_dynamics(TORQUE_WEIGHT)
    _check_cost_plugin(sc, g, eng)
    plain = _engine(sc).execute(g["parameters"][None], 2)[0][0]
    assert (g["costs"][:, 1] - plain).min() > 1e-3, "fixture's torque term is not active"
    gi = rg.load("ref_iter_tiny_torque")
    sc2 = rg.scenario("tiny", gi)
    eng2 = _engine(sc2, keep_intermediates=1)
    eng2.set_dynamics(TORQUE_WEIGHT)
    assert len(list(_check_iterations(sc2, gi, eng2))) == 3
    eng3 = _engine(sc2)                                    # without the parity taps the engine keeps its own projected copy
    eng3.set_dynamics(TORQUE_WEIGHT)
    for it in rg.iterations(gi):
        eng3.inject_noise(gi["it%d_noise" % it][None])
        cost, _, _ = eng3.iterate(it)
        assert_close(cost[0], float(gi["it%d_noiseless_cost" % it]), RTOL_F64, "noise-less cost with the torque term")
        assert_close(eng3.get(_abi.FIELD_THETA)[0], gi["it%d_theta" % it], RTOL_F64, "theta")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_engine_forward_kinematics_of_random_trees_matches_the_compiled_solvers(seed):
    g = rg.load("ref_cost_tree%d" % seed)
    sc = rg.tree_scene(seed, g)
    _check_cost_plugin(sc, g, _engine(sc))


@pytest.mark.parametrize("stem,name", [("ref_optimize_tiny_s8", "tiny"), ("ref_optimize_tiny_s7", "tiny"), ("ref_optimize_c1_s8", "C1")])
def test_engine_optimize_matches_the_compiled_reference(stem, name):
    """StompOptimizer::optimize end to end on the reference's noise: success iteration, early exit, cost log, best trajectory
    (the engine's own device-side optimize() is checked against the same host bookkeeping in test_gpu_parity.py)."""
    g = rg.load(stem)
    sc = rg.scenario(name, g, seed=int(g["seed"]))
    eng = _engine(sc)

    def step(it):
        eng.inject_noise(g["it%d_noise" % it][None])
        cost, cf, _ = eng.iterate(it)
        return float(cost[0]), int(cf[0]), 1, eng.get(_abi.FIELD_NOISELESS_TRAJECTORY)[0].copy()

    res = rg.host_bookkeeping(step, int(g["max_iterations"]), int(g["max_iterations_after_collision_free"]))
    assert res["success"] == bool(g["stats_success"])
    assert res["success_iteration"] == int(g["stats_success_iteration"]) == int(g["stats_collision_success_iteration"])
    assert res["iterations"] == int(g["stats_iterations"])
    assert res["last_improvement_iteration"] == int(g["stats_last_improvement_iteration"])
    assert_close(res["costs"], g["stats_costs"], RTOL_F64, "STOMPStatistics.costs")
    assert_close(res["best_cost"], float(g["stats_best_cost"]), RTOL_F64, "best cost")
    assert_close(res["best_trajectory"], g["stats_best_trajectory"], RTOL_F64, "best group trajectory")


def test_engine_distance_field_occupancy_matches_the_compiled_collision_space():
    """stomp_engine_build_sdf's zero set = the cells StompCollisionSpace::addCollisionObjectsToPoints marks (bit-exact)."""
    from stomp_motion_planner_icra2011_b200 import scenes
    from tests.golden.make_ref_golden import SDF_SCENE, map_points, sdf_objects
    boxes, cyls = sdf_objects()
    eng = _engine(scenes.make_scenario("tiny", num_problems=1))
    for stem, pts in (("ref_collision_cells", None), ("ref_collision_cells_points", map_points())):   # + a collision map
        g = rg.load(stem)
        eng.build_sdf(boxes=boxes, cylinders=cyls, max_distance=0.17, points=pts, **SDF_SCENE)
        got, dtype = eng.get_sdf()
        want = np.unpackbits(g["occupancy_bits"])[:got.size].reshape(g["shape"]).astype(bool)
        assert dtype == _abi.VOXEL_U8_SQ and got.shape == want.shape
        np.testing.assert_array_equal(got == 0, want)
