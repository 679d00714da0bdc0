"""Pins the CPU oracle to OUTPUTS OF THE REFERENCE ITSELF for the PI^2 half of the path.

tests/golden/ref_pi2_*.npz were produced by the reference's own, unmodified translation units
(src/policy_improvement.cpp, src/policy_improvement_loop.cpp, src/covariant_trajectory_policy.cpp, src/stomp_cost.cpp,
include/.../multivariate_gaussian.h) compiled against stand-in Eigen 2 / roscpp / Boost headers (oracle/ref_shim/,
oracle/ref_driver.cpp, tests/golden/make_ref_golden.py).  The oracle is fed the noise the reference drew
(host-injection mode) and must reproduce every intermediate of PolicyImprovementLoop::runSingleIteration.

Tolerance 1e-8 relative to the array's largest magnitude: the oracle and the compiled reference sum dense products in
different orders and invert R (cond ~ 6e6) with different eliminations; observed agreement is 1e-13 (N=20) to 3e-9 (N=100).
Not pinned by these vectors: the cost plugin (StompOptimizer::execute cannot be compiled here) — `state_costs` in the
fixtures are the oracle's own answers to the reference's Task::execute calls.
"""
import os

import numpy as np
import pytest

from oracle import oracle, reference_pi2
from stomp_motion_planner_icra2011_b200 import _abi, scenes
from tests.helpers import assert_close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("tiny", 0), ("tiny", 1), ("C1", 1), ("C1", 0)]
RTOL = 1e-8

FIELDS = ((_abi.FIELD_NOISE_PROJECTED, "noise_projected"), (_abi.FIELD_PARAMETERS, "parameters"),
          (_abi.FIELD_STATE_COSTS, "state_costs"), (_abi.FIELD_CONTROL_COSTS, "control_costs"),
          (_abi.FIELD_CUMULATIVE_COSTS, "cumulative_costs"), (_abi.FIELD_PROBABILITIES, "probabilities"),
          (_abi.FIELD_UPDATES, "updates"), (_abi.FIELD_THETA, "theta"), (_abi.FIELD_ROLLOUT_TOTAL_COSTS, "totals"))


def _load(name, cumulative):
    return np.load(os.path.join(GOLDEN, "ref_pi2_%s_c%d.npz" % (name.lower(), cumulative)))


def _iterations(g):
    return sorted(int(k[2:k.index("_")]) for k in g.files if k.endswith("_theta"))


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_setup_matrices_match_the_compiled_reference(name):
    """R, R^-1, M (column-max scaling), chol(R^-1), the joint-limit Q^-1 and the min-control-cost trajectory."""
    g = _load(name, 1)
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=1)
    o = oracle.Oracle(sc, 0)
    for f, nm in ((_abi.FIELD_CONTROL_COST, "control_cost_matrix"), (_abi.FIELD_INV_CONTROL_COST, "inv_control_cost_matrix"),
                  (_abi.FIELD_PROJECTION, "projection_matrix"), (_abi.FIELD_NOISE_CHOLESKY, "covariance_cholesky"),
                  (_abi.FIELD_QUAD_COST_INV, "quad_cost_inv"), (_abi.FIELD_THETA, "theta0")):
        assert_close(o.get(f), g[nm], 1e-7 if nm == "covariance_cholesky" else RTOL, nm)
    # KAT from src/policy_improvement.cpp:430-436, now checked on the reference's own M
    np.testing.assert_allclose(g["projection_matrix"].max(axis=0), 1.0 / sc.num_time_steps, rtol=1e-12)
    assert float(g["movement_dt"][0]) == sc.movement_duration / (sc.num_time_steps + 1)


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_control_cost_function_matches_the_compiled_reference(name):
    g = _load(name, 1)
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=1)
    o = oracle.Oracle(sc, 0)
    got = o.compute_control_costs(g["cc_parameters"][None], g["cc_noise"][None], 0.5 * sc.smoothness_cost_weight)[0]
    assert_close(got, g["cc_out"], RTOL, "computeControlCosts")


@pytest.mark.parametrize("name,cumulative", CASES)
def test_oracle_reproduces_the_reference_iterations(name, cumulative):
    g = _load(name, cumulative)
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=cumulative)
    o = oracle.Oracle(sc, 0)
    assert_close(o.get(_abi.FIELD_THETA), g["theta0"], RTOL, "theta0")
    for it in _iterations(g):
        k = "it%d_" % it
        _, _, ngen = o.iterate(it, g[k + "noise"])
        assert ngen == int(g[k + "num_rollouts_gen"])
        np.testing.assert_array_equal(o.get(_abi.FIELD_NOISE)[:ngen], g[k + "noise"])
        for f, nm in FIELDS:
            assert_close(o.get(f), g[k + nm], RTOL, "%s it %d" % (nm, it))
        # the reuse order is integer work: the ranking of (getCost(), index) must be the reference's
        assert np.array_equal(np.argsort(o.get(_abi.FIELD_ROLLOUT_TOTAL_COSTS), kind="stable"),
                              np.argsort(g[k + "totals"], kind="stable"))


def test_reused_rollouts_are_the_references_choice():
    """slots [R_gen, R) of iteration i+1 hold the R_reuse cheapest of iteration i's R rollouts + the noise-less one."""
    g = _load("C1", 1)
    R, Rre = 10, 5
    for it in _iterations(g)[:-1]:
        tot = g["it%d_totals" % it]                      # [R] + extra
        order = sorted(range(R + 1), key=lambda r: (tot[r], -1 if r == R else r))[:Rre]
        prev_params = np.concatenate([g["it%d_parameters" % it], g["it%d_theta" % it][None]])
        np.testing.assert_array_equal(g["it%d_parameters" % (it + 1)][R - Rre:], prev_params[order])


@pytest.mark.skipif(not os.path.isdir(reference_pi2.REFERENCE_ROOT), reason="/root/reference is not on this machine")
def test_fixtures_are_what_the_compiled_reference_produces_today():
    """Rebuild-and-rerun check (only where the reference sources exist): the committed vectors are reproducible."""
    from tests.golden import make_ref_golden
    fresh = make_ref_golden.run_case("tiny", 1, 5)
    g = _load("tiny", 1)
    assert sorted(fresh) == sorted(g.files)
    for k in g.files:
        np.testing.assert_allclose(fresh[k], g[k], rtol=1e-12, atol=1e-300, err_msg=k)
