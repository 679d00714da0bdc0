"""ctypes binding of oracle/_ref/libstomp_ref_pi2.so: the reference's OWN PolicyImprovementLoop /
PolicyImprovement / CovariantTrajectoryPolicy / MultivariateGaussian / StompCost translation units, compiled
unmodified from /root/reference against the stand-in headers in oracle/ref_shim/ (see oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY.  It can only be built where /root/reference exists (the authoring container);
tests/golden/make_ref_golden.py runs it there and commits the vectors, which is what travels to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libstomp_ref_pi2.so")
REFERENCE_ROOT = "/root/reference/stomp_motion_planner"
_lib = None

EXECUTE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)


def available():
    """True when the compiled reference exists or can be built here."""
    return os.path.exists(_LIB_PATH) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
            raise RuntimeError("reference sources are not present; oracle/_ref cannot be built on this machine")
        subprocess.check_call(["make", "-C", _HERE, "ref"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.stomp_ref_create.restype = C.c_void_p
        _lib.stomp_ref_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          EXECUTE_CB, C.c_void_p]
        _lib.stomp_ref_destroy.argtypes = [C.c_void_p]
        _lib.stomp_ref_run_single_iteration.argtypes = [C.c_void_p, C.c_int]
        _lib.stomp_ref_set_parameters.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _lib.stomp_ref_get_parameters.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _lib.stomp_ref_compute_control_costs.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                         C.c_double, C.POINTER(C.c_double)]
        _lib.stomp_ref_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
        _lib.stomp_ref_quad_cost_inv.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_int,
                                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class ReferencePI2:
    """The reference's PolicyImprovementLoop around a Python cost plugin `execute(parameters[D][N], iteration) -> costs[N]`."""

    def __init__(self, N, D, R, R_reuse, movement_duration, ridge, derivative_costs, noise_stddev, noise_decay,
                 control_cost_weight, use_cumulative_costs, start, goal, execute):
        self.N, self.D, self.R = N, D, R
        self.L = lib()
        self._execute = execute
        self.calls = []  # (iteration_number, parameters[D][N], costs[N]) of every Task::execute call, in order

        def cb(_user, params, costs, iteration_number):
            p = np.ctypeslib.as_array(params, shape=(D, N)).copy()
            c = np.asarray(self._execute(p, iteration_number), dtype=np.float64)
            self.calls.append((iteration_number, p, c.copy()))
            np.ctypeslib.as_array(costs, shape=(N,))[:] = c
            return 0

        self._cb = EXECUTE_CB(cb)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self._keep = [f(derivative_costs), f(noise_stddev), f(noise_decay), f(start), f(goal)]
        k = self._keep
        self.h = self.L.stomp_ref_create(N, D, R, R_reuse, float(movement_duration), float(ridge), _dp(k[0]), _dp(k[1]),
                                         _dp(k[2]), float(control_cost_weight), int(use_cumulative_costs), _dp(k[3]),
                                         _dp(k[4]), self._cb, None)
        if not self.h:
            raise RuntimeError("stomp_ref_create failed")

    def close(self):
        if self.h:
            self.L.stomp_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_single_iteration(self, iteration_number):
        if self.L.stomp_ref_run_single_iteration(self.h, iteration_number) != 0:
            raise RuntimeError("runSingleIteration returned false")

    def get_parameters(self):
        out = np.empty((self.D, self.N))
        self.L.stomp_ref_get_parameters(self.h, _dp(out))
        return out

    def set_parameters(self, theta):
        t = np.ascontiguousarray(theta, dtype=np.float64)
        self.L.stomp_ref_set_parameters(self.h, _dp(t))

    def compute_control_costs(self, parameters, noise, weight):
        p, e = np.ascontiguousarray(parameters, dtype=np.float64), np.ascontiguousarray(noise, dtype=np.float64)
        out = np.empty((self.D, self.N))
        if self.L.stomp_ref_compute_control_costs(self.h, _dp(p), _dp(e), float(weight), _dp(out)) != 0:
            raise RuntimeError("computeControlCosts returned false")
        return out

    def get(self, field):
        R, D, N = self.R, self.D, self.N
        base = field[6:] if field.startswith("extra_") else field
        n_r = 1 if field.startswith("extra_") else R
        shapes = {"state_costs": (n_r, N), "total": (n_r,), "control_cost_matrix": (N, N), "inv_control_cost_matrix": (N, N),
                  "projection_matrix": (N, N), "covariance_cholesky": (N, N), "control_cost_matrix_all": (N + 12, N + 12),
                  "parameters_all": (D, N + 12), "parameter_updates": (D, N), "num_rollouts_gen": (1,), "movement_dt": (1,)}
        out = np.empty(shapes.get(base, (n_r, D, N)))
        rc = self.L.stomp_ref_get(self.h, field.encode(), _dp(out))
        if rc != 0:
            raise KeyError(field)
        return out


def quad_cost_inv(num_vars_all, discretization, smoothness_costs, ridge, joint_cost):
    """StompCost::getQuadraticCostInverse() per joint after the optimizer's global scaling (src/stomp_optimizer.cpp:104-125)."""
    jc = np.ascontiguousarray(joint_cost, dtype=np.float64)
    sc = np.ascontiguousarray(smoothness_costs, dtype=np.float64)
    n = num_vars_all - 12
    out = np.empty((len(jc), n, n))
    lib().stomp_ref_quad_cost_inv(num_vars_all, float(discretization), _dp(sc), float(ridge), len(jc), _dp(jc), _dp(out))
    return out
