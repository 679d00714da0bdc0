"""Shared helpers of the parity tests: seeded noise drawn on the host exactly the way the reference's
MultivariateGaussian does (eps = sigma * chol(R^-1) * z), so that the oracle and the engine receive
identical injected noise (BASELINE.json north_star: host-injection mode)."""
import contextlib
import os

import numpy as np

from stomp_motion_planner_icra2011_b200 import _abi

# north_star tolerances
RTOL_F64 = 1e-5
RTOL_F32 = 1e-3


def correlated_noise(L, rng, shape_prefix, sigma):
    """shape_prefix + (D, N) noise; L = lower Cholesky factor of R^-1 [N][N]; sigma [D]."""
    D, N = len(sigma), L.shape[0]
    z = rng.standard_normal(tuple(shape_prefix) + (D, N))
    return np.einsum("ij,...j->...i", L, z) * np.asarray(sigma)[:, None]


def assert_close(actual, desired, rtol, what, atol_scale=1e-9):
    """relative tolerance `rtol` against the largest magnitude of the reference array (so that entries that
    are exactly or nearly zero in the reference are compared on the array's own scale)."""
    actual, desired = np.asarray(actual), np.asarray(desired)
    assert actual.shape == desired.shape, (what, actual.shape, desired.shape)
    scale = float(np.max(np.abs(desired))) if desired.size else 0.0
    np.testing.assert_allclose(actual, desired, rtol=rtol, atol=atol_scale * scale + 1e-20, err_msg=what)   # 1e-20: rounding dust of exactly-zero quantities


def oracle_batch(oracle_mod, sc):
    return [oracle_mod.Oracle(sc, b) for b in range(sc.start.shape[0])]


@contextlib.contextmanager
def band_solves_only():
    """Engines created inside always use k_generate's band solves (STOMP_NO_DENSE is read at stomp_engine_create); small
    batches otherwise take k_generate_dense, and the bench-size path would go untested by the small parity cases."""
    old = os.environ.get("STOMP_NO_DENSE")
    os.environ["STOMP_NO_DENSE"] = "1"
    try:
        yield
    finally:
        if old is None:
            del os.environ["STOMP_NO_DENSE"]
        else:
            os.environ["STOMP_NO_DENSE"] = old
