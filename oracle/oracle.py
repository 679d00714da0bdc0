"""ctypes binding of the CPU oracle (oracle/stomp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from stomp_motion_planner_icra2011_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstomp_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "stomp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libstomp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.stomp_oracle_last_error.restype = C.c_char_p
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class Oracle:
    """One planning problem (the reference plans one request at a time)."""

    def __init__(self, scenario, problem=0):
        self.sc = scenario
        self.L = lib()
        self.D, self.N, self.R = scenario.robot.num_dimensions, scenario.num_time_steps, scenario.num_rollouts
        self.K = len(scenario.robot.spheres)
        desc = scenario.desc(num_problems=1)
        self.h = C.c_void_p()
        self._ck(self.L.stomp_oracle_create(C.byref(desc), C.byref(self.h)))
        rb = scenario.robot
        segs, sph, lim = rb.c_segments(), rb.c_spheres(), rb.c_limits()
        self._ck(self.L.stomp_oracle_set_robot(self.h, segs, len(rb.segments), rb.reference_segment, sph,
                                               len(rb.spheres), lim))
        sdf = scenario.sdf
        org = (C.c_double * 3)(*sdf.origin)
        nx, ny, nz = sdf.dims
        self._ck(self.L.stomp_oracle_set_sdf(self.h, sdf.voxels.ctypes.data_as(C.c_void_p), nx, ny, nz, org,
                                             C.c_double(sdf.resolution), sdf.voxel_dtype))
        self._ck(self.L.stomp_oracle_set_noise(self.h, _dp(np.ascontiguousarray(scenario.noise_stddev, dtype=np.float64)),
                                               _dp(np.ascontiguousarray(scenario.noise_decay, dtype=np.float64))))
        self.set_problem(scenario.start[problem], scenario.goal[problem])

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.stomp_oracle_last_error().decode())

    def close(self):
        if self.h:
            self.L.stomp_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_problem(self, start, goal):
        s = np.ascontiguousarray(start, dtype=np.float64)
        g = np.ascontiguousarray(goal, dtype=np.float64)
        self._ck(self.L.stomp_oracle_set_problem(self.h, _dp(s), _dp(g)))

    def set_constraints(self, constraints, weight):
        arr = (_abi.OrientationConstraint * max(1, len(constraints)))()
        for i, c in enumerate(constraints):
            arr[i].segment, arr[i].body_fixed = c["segment"], int(c.get("body_fixed", 0))
            arr[i].orientation[:] = c["orientation"]
            arr[i].absolute_roll_tolerance, arr[i].absolute_pitch_tolerance, arr[i].absolute_yaw_tolerance = c["tolerances"]
            arr[i].weight = c.get("weight", 1.0)
        self._ck(self.L.stomp_oracle_set_constraints(self.h, arr, len(constraints), C.c_double(weight)))

    def set_dynamics(self, torque_cost_weight, gravity=(0.0, 0.0, -9.8)):
        """torque term of StompOptimizer::execute over the robot's inverse-dynamics chain (Robot.inertias / Robot.chain)."""
        rb = self.sc.robot
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        self._ck(self.L.stomp_oracle_set_dynamics(self.h, rb.c_inertias(), rb.chain[0], rb.chain[1], _dp(g),
                                                  C.c_double(torque_cost_weight)))

    def last_torques(self):
        out = np.empty((self.N, self.D))
        self._ck(self.L.stomp_oracle_last_torques(self.h, _dp(out)))
        return out

    def execute_constraints_satisfied(self, n):
        out = np.empty(n, dtype=np.int32)
        self._ck(self.L.stomp_oracle_execute_constraints_satisfied(self.h, _ip(out), C.c_size_t(n)))
        return out

    def seed(self, seed):
        self._ck(self.L.stomp_oracle_seed(self.h, C.c_uint64(seed)))

    def set_parameters(self, theta):
        t = np.ascontiguousarray(theta, dtype=np.float64)
        self._ck(self.L.stomp_oracle_set_parameters(self.h, _dp(t)))

    def get_parameters(self):
        out = np.empty((self.D, self.N))
        self._ck(self.L.stomp_oracle_get_parameters(self.h, _dp(out)))
        return out

    def update_parameters(self, updates):
        u = np.ascontiguousarray(updates, dtype=np.float64)
        self._ck(self.L.stomp_oracle_update_parameters(self.h, _dp(u)))

    def compute_control_costs(self, parameters, noise, weight):
        p = np.ascontiguousarray(parameters, dtype=np.float64)
        e = np.ascontiguousarray(noise, dtype=np.float64)
        n = p.shape[0]
        out = np.empty((n, self.D, self.N))
        self._ck(self.L.stomp_oracle_compute_control_costs(self.h, _dp(p), _dp(e), n, C.c_double(weight), _dp(out)))
        return out

    def execute(self, parameters, iteration_number=2):
        p = np.ascontiguousarray(parameters, dtype=np.float64).reshape(-1, self.D, self.N)
        n = p.shape[0]
        costs = np.empty((n, self.N))
        cf = np.empty(n, dtype=np.int32)
        self._ck(self.L.stomp_oracle_execute(self.h, _dp(p), n, iteration_number, _dp(costs), _ip(cf)))
        return costs, cf

    def execute_debug(self, parameters):
        p = np.ascontiguousarray(parameters, dtype=np.float64).reshape(self.D, self.N)
        dbg = (_abi.SphereDebug * ((self.N + 3) * self.K))()
        clipped = np.empty((self.D, self.N))
        self._ck(self.L.stomp_oracle_execute_debug(self.h, _dp(p), dbg, _dp(clipped)))
        return debug_to_arrays(dbg, self.N, self.K), clipped

    def get_rollouts(self, noise_stddev, eps=None):
        ns = np.ascontiguousarray(noise_stddev, dtype=np.float64)
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        out = np.empty((self.R, self.D, self.N))
        ngen = C.c_int32()
        self._ck(self.L.stomp_oracle_get_rollouts(self.h, _dp(ns), _dp(e), _dp(out), C.byref(ngen)))
        return out[:ngen.value].copy()

    def set_rollout_costs(self, costs, control_cost_weight):
        c = np.ascontiguousarray(costs, dtype=np.float64)
        totals = np.empty(self.R)
        self._ck(self.L.stomp_oracle_set_rollout_costs(self.h, _dp(c), C.c_double(control_cost_weight), _dp(totals)))
        return totals

    def improve_policy(self):
        out = np.empty((self.D, self.N))
        self._ck(self.L.stomp_oracle_improve_policy(self.h, _dp(out)))
        return out

    def add_extra_rollouts(self, costs):
        c = np.ascontiguousarray(costs, dtype=np.float64)
        self._ck(self.L.stomp_oracle_add_extra_rollouts(self.h, _dp(c)))

    def iterate(self, iteration_number, eps=None):
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        cost, cf, ngen = C.c_double(), C.c_int32(), C.c_int32()
        self._ck(self.L.stomp_oracle_iterate(self.h, iteration_number, _dp(e), C.byref(cost), C.byref(cf), C.byref(ngen)))
        return cost.value, cf.value, ngen.value

    def get(self, field):
        R, D, N = self.R, self.D, self.N
        shapes = {
            _abi.FIELD_THETA: (D, N), _abi.FIELD_NOISE: (R, D, N), _abi.FIELD_PARAMETERS: (R, D, N),
            _abi.FIELD_NOISE_PROJECTED: (R, D, N), _abi.FIELD_STATE_COSTS: (R, N), _abi.FIELD_CONTROL_COSTS: (R, D, N),
            _abi.FIELD_CUMULATIVE_COSTS: (R, D, N), _abi.FIELD_PROBABILITIES: (R, D, N), _abi.FIELD_UPDATES: (D, N),
            _abi.FIELD_NOISELESS_COSTS: (N,), _abi.FIELD_ROLLOUT_TOTAL_COSTS: (R + 1,),
            _abi.FIELD_INV_CONTROL_COST: (N, N), _abi.FIELD_NOISE_CHOLESKY: (N, N), _abi.FIELD_PROJECTION: (N, N),
            _abi.FIELD_QUAD_COST_INV: (N, N), _abi.FIELD_CONTROL_COST: (N, N),
        }
        if field in (_abi.FIELD_COLLISION_FREE, _abi.FIELD_CONSTRAINTS_SATISFIED):
            out = np.empty(R + 1, dtype=np.int32)
        else:
            out = np.empty(shapes[field])
        self._ck(self.L.stomp_oracle_get(self.h, field, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out


def debug_to_arrays(dbg, N, K):
    a = np.frombuffer(dbg, dtype=np.dtype([("voxel", np.int32, 3), ("in_collision", np.int32), ("position", np.float64, 3),
                                           ("potential", np.float64), ("vel_mag", np.float64)])).reshape(N + 3, K)
    return {k: a[k].copy() for k in a.dtype.names}
