"""ctypes binding of oracle/_ref/libstomp_ref.so: the reference's OWN translation units (PolicyImprovementLoop,
PolicyImprovement, CovariantTrajectoryPolicy, MultivariateGaussian, StompCost, StompOptimizer, StompTrajectory,
StompCollisionPoint, TreeFkSolverJointPosAxis[Partial], OrientationConstraintEvaluator, StompParameters) compiled
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
_LIB_PATH = os.path.join(_HERE, "_ref", "libstomp_ref.so")
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


class ReferenceOptimizer:
    """The reference's OWN StompOptimizer (cost plugin, joint limits, FK, collision potential, optimize() loop) compiled
    unmodified (oracle/ref_driver.cpp, second half), set up from the same tables the engine and the oracle receive."""

    def __init__(self, scenario, problem=0, constraints=(), constraint_cost_weight=0.0, max_iterations=500,
                 max_iterations_after_collision_free=100):
        from stomp_motion_planner_icra2011_b200 import _abi
        from oracle.oracle import debug_to_arrays
        self._abi, self._debug_to_arrays = _abi, debug_to_arrays
        sc = self.sc = scenario
        self.L = lib()
        L = self.L
        L.stomp_ref_opt_create.restype = C.c_void_p
        for fn in ("stomp_ref_opt_destroy", "stomp_ref_opt_info", "stomp_ref_opt_execute", "stomp_ref_opt_debug",
                   "stomp_ref_opt_frames", "stomp_ref_opt_begin", "stomp_ref_opt_iterate", "stomp_ref_opt_get",
                   "stomp_ref_opt_optimize"):
            getattr(L, fn).restype = C.c_int
        rb = sc.robot
        self.D, self.N, self.R = rb.num_dimensions, sc.num_time_steps, sc.num_rollouts
        self.K, self.S = len(rb.spheres), len(rb.segments)
        self.max_iterations = max_iterations
        desc = sc.desc(num_problems=1)
        sdf = sc.sdf
        nx, ny, nz = sdf.dims
        cons = (_abi.OrientationConstraint * max(1, len(constraints)))()
        for i, c in enumerate(constraints):
            cons[i].segment, cons[i].body_fixed = c["segment"], int(c.get("body_fixed", 0))
            cons[i].orientation[:] = c["orientation"]
            cons[i].absolute_roll_tolerance, cons[i].absolute_pitch_tolerance, cons[i].absolute_yaw_tolerance = c["tolerances"]
            cons[i].weight = c.get("weight", 1.0)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        k = self._keep = [f(sdf.origin), f(sc.noise_stddev), f(sc.noise_decay), f(sc.start[problem]), f(sc.goal[problem])]
        self.h = C.c_void_p(L.stomp_ref_opt_create(
            C.byref(desc), rb.c_segments(), self.S, rb.reference_segment, rb.c_spheres(), self.K, rb.c_limits(),
            sdf.voxels.ctypes.data_as(C.c_void_p), nx, ny, nz, _dp(k[0]), C.c_double(sdf.resolution), sdf.voxel_dtype,
            _dp(k[1]), _dp(k[2]), _dp(k[3]), _dp(k[4]), cons, len(constraints), C.c_double(constraint_cost_weight),
            int(max_iterations), int(max_iterations_after_collision_free)))
        if not self.h:
            raise RuntimeError("stomp_ref_opt_create failed")
        self.info = np.empty(7)
        L.stomp_ref_opt_info(self.h, _dp(self.info))
        assert int(self.info[4]) == self.S and int(self.info[1]) == self.N
        # the segment table must be in DFS pre-order, the order kdl_parser builds trees in: KDL renumbers joints in that
        # order whenever a Tree is copied (the FK solvers keep a copy), so any other order is outside the reference's domain
        if self.info[5] != 1.0:
            raise ValueError("segment table is not in DFS pre-order")
        # the reference derives the policy's movement duration itself (int-truncated group-trajectory duration)
        self.movement_duration = float(self.info[0])

    def close(self):
        if self.h:
            self.L.stomp_ref_opt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_dynamics(self, torque_cost_weight, gravity=(0.0, 0.0, -9.8)):
        rb = self.sc.robot
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        rc = self.L.stomp_ref_opt_set_dynamics(self.h, rb.c_inertias(), rb.chain[0], rb.chain[1], _dp(g),
                                               C.c_double(torque_cost_weight))
        if rc != 0:
            raise RuntimeError("stomp_ref_opt_set_dynamics failed (rc %d)" % rc)

    def execute(self, parameters, iteration_number=2):
        """StompOptimizer::execute -> (costs[N], collision_free, constraints_satisfied)."""
        p = np.ascontiguousarray(parameters, dtype=np.float64).reshape(self.D, self.N)
        costs = np.empty(self.N)
        cf, cs = C.c_int32(), C.c_int32()
        if self.L.stomp_ref_opt_execute(self.h, _dp(p), int(iteration_number), _dp(costs), C.byref(cf), C.byref(cs)) != 0:
            raise RuntimeError("execute returned false")
        return costs, cf.value, cs.value

    def execute_debug(self, parameters):
        """same layout as Oracle.execute_debug: execute at iteration 1, per-sphere records of points -1..N+1 + clipped trajectory."""
        costs, cf, _ = self.execute(parameters, 1)
        dbg = (self._abi.SphereDebug * ((self.N + 3) * self.K))()
        clipped = np.empty((self.D, self.N))
        self.L.stomp_ref_opt_debug(self.h, dbg, _dp(clipped))
        out = self._debug_to_arrays(dbg, self.N, self.K)
        out["costs"], out["collision_free"] = costs, cf
        return out, clipped

    def frames(self, t):
        out = np.empty((self.S, 12))
        self.L.stomp_ref_opt_frames(self.h, int(t), _dp(out))
        return out[:, :9].reshape(self.S, 3, 3), out[:, 9:]

    def begin(self):
        if self.L.stomp_ref_opt_begin(self.h) != 0:
            raise RuntimeError("PolicyImprovementLoop::initialize failed")

    def iterate(self, iteration_number):
        cost, cf, cs = C.c_double(), C.c_int32(), C.c_int32()
        if self.L.stomp_ref_opt_iterate(self.h, int(iteration_number), C.byref(cost), C.byref(cf), C.byref(cs)) != 0:
            raise RuntimeError("runSingleIteration returned false")
        return cost.value, cf.value, cs.value

    def get(self, field):
        R, D, N = self.R, self.D, self.N
        if field in ("exec_parameters", "exec_costs", "exec_collision_free"):
            n = int(self.get("exec_count")[0])
            out = np.empty({"exec_parameters": (n, D, N), "exec_costs": (n, N), "exec_collision_free": (n,)}[field])
        else:
            base = field[6:] if field.startswith("extra_") else field
            n_r = 1 if field.startswith("extra_") else R
            shapes = {"state_costs": (n_r, N), "total": (n_r,), "control_cost_matrix": (N, N), "inv_control_cost_matrix": (N, N),
                      "projection_matrix": (N, N), "covariance_cholesky": (N, N), "quad_cost_inv": (N, N),
                      "parameters_all": (D, N + 12), "parameter_updates": (D, N), "num_rollouts_gen": (1,), "movement_dt": (1,),
                      "exec_count": (1,), "exec_clear": (1,), "best_group_trajectory": (D, N), "theta": (D, N)}
            out = np.empty(shapes.get(base, (n_r, D, N)))
        rc = self.L.stomp_ref_opt_get(self.h, field.encode(), _dp(out))
        if rc != 0:
            raise KeyError("%s (rc %d)" % (field, rc))
        return out

    def optimize(self):
        """the reference's whole StompOptimizer::optimize(); returns the STOMPStatistics it publishes + the rollouts it executed."""
        stats, costs = np.empty(6), np.zeros(self.max_iterations)
        if self.L.stomp_ref_opt_optimize(self.h, _dp(stats), _dp(costs)) != 0:
            raise RuntimeError("optimize published no statistics")
        n = int(stats[4])
        return dict(success=bool(stats[0]), success_iteration=int(stats[1]), collision_success_iteration=int(stats[2]),
                    best_cost=float(stats[3]), iterations=n, last_improvement_iteration=int(stats[5]), costs=costs[:n].copy(),
                    best_trajectory=self.get("best_group_trajectory"))


def collision_points(robot, links, default_clearance=0.07, attached=(), attached_padding=0.0):
    """StompRobotModel::generateLinkCollisionPoints + generateAttachedObjectCollisionPoints + populatePlanningGroupCollisionPoints
    on the robot's segment table.  links: [(segment, radius, clearance | None, extension)] in getGroupLinkUnion() order;
    attached: [(segment, 'sphere' | 'box' | 'cylinder', dims, position in the link frame)].
    Returns [(segment, radius, clearance, (x, y, z))] = the planning group's collision points in the reference's order."""
    from stomp_motion_planner_icra2011_b200 import _abi
    L = lib()
    n = len(links)
    seg = np.array([l[0] for l in links], dtype=np.int32)
    rad = np.array([l[1] for l in links], dtype=np.float64)
    clr = np.array([-1.0 if l[2] is None else l[2] for l in links], dtype=np.float64)
    ext = np.array([l[3] for l in links], dtype=np.float64)
    shape_id = {"sphere": 0, "box": 1, "cylinder": 2}
    aseg = np.array([a[0] for a in attached] + [0], dtype=np.int32)
    ash = np.array([shape_id[a[1]] for a in attached] + [0], dtype=np.int32)
    adims = np.zeros((len(attached) + 1, 3))
    apos = np.zeros((len(attached) + 1, 3))
    for i, a in enumerate(attached):
        adims[i, :len(a[2])] = a[2]
        apos[i] = a[3]
    out = (_abi.Sphere * 4096)()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    L.stomp_ref_collision_points.restype = C.c_int
    cnt = L.stomp_ref_collision_points(robot.c_segments(), len(robot.segments), robot.reference_segment, robot.num_dimensions,
                                       ip(seg), _dp(rad), _dp(clr), _dp(ext), n, C.c_double(default_clearance), ip(aseg), ip(ash),
                                       _dp(adims), _dp(apos), len(attached), C.c_double(attached_padding), out, 4096)
    if cnt < 0 or cnt > 4096:
        raise RuntimeError("stomp_ref_collision_points failed (%d)" % cnt)
    return [(out[i].segment, out[i].radius, out[i].clearance, tuple(out[i].pos)) for i in range(cnt)]


def collision_object_cells(size, origin, resolution, boxes=(), cylinders=(), points=None):
    """StompCollisionSpace::addCollisionObjectsToPoints + the distance field's cell binning -> (occupancy[nx][ny][nz] bool,
    number of lattice points generated).  boxes: [(position, quaternion xyzw, dimensions)], cylinders: [(position, quaternion,
    radius, height)]."""
    from stomp_motion_planner_icra2011_b200 import _abi
    L = lib()
    barr = (_abi.Box * max(1, len(boxes)))()
    for i, (p, q, d) in enumerate(boxes):
        barr[i].position[:], barr[i].orientation[:], barr[i].dimensions[:] = p, q, d
    carr = (_abi.Cylinder * max(1, len(cylinders)))()
    for i, (p, q, r, h) in enumerate(cylinders):
        carr[i].position[:], carr[i].orientation[:] = p, q
        carr[i].radius, carr[i].height = r, h
    sz, org = np.ascontiguousarray(size, dtype=np.float64), np.ascontiguousarray(origin, dtype=np.float64)
    n = [int(size[i] / resolution) for i in range(3)]
    occ = np.zeros(n, dtype=np.uint8)
    dims = (C.c_int32 * 3)()
    L.stomp_ref_collision_object_cells.restype = C.c_longlong
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3) if points is not None else np.zeros((0, 3))
    cnt = L.stomp_ref_collision_object_cells(_dp(sz), _dp(org), C.c_double(resolution), barr, len(boxes), carr, len(cylinders),
                                             _dp(pts), C.c_int64(len(pts)), occ.ctypes.data_as(C.POINTER(C.c_uint8)), dims)
    assert list(dims) == n, (list(dims), n)
    return occ.astype(bool), int(cnt)
