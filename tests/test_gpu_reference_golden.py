"""GPU parity against OUTPUTS OF THE REFERENCE ITSELF (tests/golden/ref_pi2_*.npz; see tests/test_reference_pinning.py
and tests/golden/make_ref_golden.py for how the reference's own PI^2 translation units were compiled and run).

The engine is fed the noise the reference's MultivariateGaussian drew (host-injection mode, BASELINE north_star) through
the C ABI and must land on the reference's intermediates and on its updated policy within 1e-5 relative (fp64 mode).
No oracle is needed at run time.  The state costs inside the fixtures are the CPU restatement's cost-plugin answers
(the reference's StompOptimizer::execute cannot be compiled here); the engine recomputes them with its own k_cost.
"""
import os

import numpy as np
import pytest

from stomp_motion_planner_icra2011_b200 import _abi, scenes
from tests.helpers import RTOL_F64, assert_close

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ((_abi.FIELD_NOISE_PROJECTED, "noise_projected"), (_abi.FIELD_PARAMETERS, "parameters"),
          (_abi.FIELD_STATE_COSTS, "state_costs"), (_abi.FIELD_CONTROL_COSTS, "control_costs"),
          (_abi.FIELD_CUMULATIVE_COSTS, "cumulative_costs"), (_abi.FIELD_PROBABILITIES, "probabilities"),
          (_abi.FIELD_UPDATES, "updates"), (_abi.FIELD_THETA, "theta"), (_abi.FIELD_ROLLOUT_TOTAL_COSTS, "totals"))


def _load(name, cumulative):
    return np.load(os.path.join(GOLDEN, "ref_pi2_%s_c%d.npz" % (name.lower(), cumulative)))


def _iterations(g):
    return sorted(int(k[2:k.index("_")]) for k in g.files if k.endswith("_theta"))


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_engine_setup_matches_the_compiled_reference(name):
    from stomp_motion_planner_icra2011_b200.engine import Engine
    g = _load(name, 1)
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=1)
    eng = Engine(sc)
    for f, nm in ((_abi.FIELD_CONTROL_COST, "control_cost_matrix"), (_abi.FIELD_INV_CONTROL_COST, "inv_control_cost_matrix"),
                  (_abi.FIELD_PROJECTION, "projection_matrix"), (_abi.FIELD_NOISE_CHOLESKY, "covariance_cholesky"),
                  (_abi.FIELD_QUAD_COST_INV, "quad_cost_inv")):
        assert_close(np.asarray(eng.get(f)).reshape(g[nm].shape), g[nm], RTOL_F64, nm)
    assert_close(eng.get(_abi.FIELD_THETA)[0], g["theta0"], RTOL_F64, "min-control-cost trajectory")
    got = eng.compute_control_costs(g["cc_parameters"][None, None], g["cc_noise"][None, None], 0.5 * sc.smoothness_cost_weight)
    assert_close(np.asarray(got).reshape(g["cc_out"].shape), g["cc_out"], RTOL_F64, "computeControlCosts")


@pytest.mark.parametrize("name,cumulative", [("tiny", 0), ("tiny", 1), ("C1", 1), ("C1", 0)])
def test_engine_reproduces_the_reference_iterations(name, cumulative):
    from stomp_motion_planner_icra2011_b200.engine import Engine
    g = _load(name, cumulative)
    sc = scenes.make_scenario(name, num_problems=1, use_cumulative_costs=cumulative)
    eng = Engine(sc, keep_intermediates=1)
    R = sc.num_rollouts
    for it in _iterations(g):
        k = "it%d_" % it
        eng.inject_noise(g[k + "noise"][None])
        cost, _, ngen = eng.iterate(it)
        assert ngen == int(g[k + "num_rollouts_gen"])
        for f, nm in FIELDS:
            assert_close(eng.get(f)[0], g[k + nm], RTOL_F64, "%s it %d" % (nm, it))
        assert_close(cost[0], g[k + "exec_costs"][-1].sum(), RTOL_F64, "noise-less rollout cost")
        # integer work: the (getCost(), index) ranking that selects next iteration's reused rollouts
        tot_e, tot_r = eng.get(_abi.FIELD_ROLLOUT_TOTAL_COSTS)[0], g[k + "totals"]
        rank = lambda t: sorted(range(R + 1), key=lambda r: (t[r], -1 if r == R else r))  # noqa: E731
        assert rank(tot_e) == rank(tot_r)
