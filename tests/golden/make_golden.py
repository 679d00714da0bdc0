"""Generates tests/golden/tiny_iterations.npz with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors and cannot be built here
(SURVEY.md §8c), so these vectors pin the oracle against *itself over time* (regression) and give the GPU
tests a fixture that does not need the oracle at run time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402
from stomp_motion_planner_icra2011_b200 import _abi, scenes  # noqa: E402
from tests.helpers import correlated_noise  # noqa: E402

ITERATIONS = 4


def main():
    out = {}
    for cumulative in (0, 1):
        sc = scenes.make_scenario("tiny", num_problems=2, use_cumulative_costs=cumulative)
        ors = [Oracle(sc, b) for b in range(2)]
        rng = np.random.default_rng(2011)
        L = ors[0].get(_abi.FIELD_NOISE_CHOLESKY)
        for it in range(1, ITERATIONS + 1):
            sigma = sc.noise_stddev * sc.noise_decay ** (it - 1)
            ngen = sc.num_rollouts if it == 1 else sc.num_rollouts - sc.num_reused_rollouts
            eps = correlated_noise(L, rng, (2, ngen), sigma)
            key = "c%d_it%d_" % (cumulative, it)
            out[key + "eps"] = eps
            res = [o.iterate(it, eps[b]) for b, o in enumerate(ors)]
            out[key + "noiseless_cost"] = np.array([r[0] for r in res])
            out[key + "collision_free"] = np.array([r[1] for r in res], dtype=np.int32)
            for f, nm in ((_abi.FIELD_THETA, "theta"), (_abi.FIELD_STATE_COSTS, "state_costs"),
                          (_abi.FIELD_PROBABILITIES, "probabilities"), (_abi.FIELD_UPDATES, "updates"),
                          (_abi.FIELD_ROLLOUT_TOTAL_COSTS, "totals")):
                out[key + nm] = np.stack([o.get(f) for o in ors])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_iterations.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
