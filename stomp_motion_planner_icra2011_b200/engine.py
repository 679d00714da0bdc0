"""ctypes binding of the CUDA engine's C ABI (include/stomp_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
fallback of any kind: if the library is missing or no CUDA device is present, construction
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STOMP_B200_LIB", os.path.join(_HERE, "libstomp_b200.so"))  # env override: A/B builds
_lib = None

EXPORTS = [
    "stomp_engine_create", "stomp_engine_destroy", "stomp_engine_last_error", "stomp_engine_abi_version",
    "stomp_engine_build_info", "stomp_engine_set_robot", "stomp_engine_set_sdf", "stomp_engine_set_noise",
    "stomp_engine_set_problems", "stomp_engine_set_parameters", "stomp_engine_get_parameters",
    "stomp_engine_update_parameters", "stomp_engine_compute_control_costs", "stomp_engine_seed",
    "stomp_engine_inject_noise", "stomp_engine_sample_noise", "stomp_engine_execute", "stomp_engine_execute_debug",
    "stomp_engine_get_rollouts", "stomp_engine_set_rollout_costs", "stomp_engine_improve_policy",
    "stomp_engine_add_extra_rollouts", "stomp_engine_iterate", "stomp_engine_run", "stomp_engine_synchronize",
    "stomp_engine_get", "stomp_engine_launch_count", "stomp_engine_stream", "stomp_engine_timer_start",
    "stomp_engine_timer_stop", "stomp_engine_shard_buffers", "stomp_engine_iterate_sharded_phase",
    "stomp_engine_set_profiling", "stomp_engine_get_profile", "stomp_engine_dump_timeline", "stomp_engine_optimize",
    "stomp_engine_build_sdf", "stomp_engine_get_sdf", "stomp_engine_inject_noise_async", "stomp_engine_last_stats",
    "stomp_engine_set_constraints", "stomp_engine_execute_constraints_satisfied",
    "stomp_engine_request_results_async", "stomp_engine_wait_results", "stomp_engine_set_dynamics",
    "stomp_engine_shard_ipc_handle", "stomp_engine_shard_open_peers", "stomp_engine_iterate_sharded_fused",
    "stomp_engine_shard_status", "stomp_engine_build_sdf_points", "stomp_engine_set_graph_mode", "stomp_engine_build_sdf_bodies", "stomp_engine_build_sdf_meshes",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA engine library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                               % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.stomp_engine_last_error.restype = C.c_char_p
        _lib.stomp_engine_build_info.restype = C.c_char_p
        _lib.stomp_engine_launch_count.restype = C.c_int64
        _lib.stomp_engine_stream.restype = C.c_void_p
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """B planning problems on one GPU.  Mirrors the reference's PolicyImprovementLoop /
    PolicyImprovement / Task surface for the per-iteration rollout loop."""

    def __init__(self, scenario, dtype=_abi.F64, device=0, keep_intermediates=0, problems=None,
                 shard_rank=0, shard_world=1):
        self.sc = scenario
        self.L = lib()
        sel = slice(None) if problems is None else problems
        self.start = _f64(scenario.start[sel])
        self.goal = _f64(scenario.goal[sel])
        self.B = self.start.shape[0]
        self.D, self.N, self.R = scenario.robot.num_dimensions, scenario.num_time_steps, scenario.num_rollouts
        self.K = len(scenario.robot.spheres)
        self.desc = scenario.desc(dtype=dtype, device=device, keep_intermediates=keep_intermediates,
                                  num_problems=self.B, shard_rank=shard_rank, shard_world=shard_world)
        self.h = C.c_void_p()
        self._ck(self.L.stomp_engine_create(C.byref(self.desc), C.byref(self.h)))
        rb = scenario.robot
        self._ck(self.L.stomp_engine_set_robot(self.h, rb.c_segments(), len(rb.segments), rb.reference_segment,
                                               rb.c_spheres(), len(rb.spheres), rb.c_limits()))
        sdf = scenario.sdf
        nx, ny, nz = sdf.dims
        self._ck(self.L.stomp_engine_set_sdf(self.h, sdf.voxels.ctypes.data_as(C.c_void_p), nx, ny, nz,
                                             (C.c_double * 3)(*sdf.origin), C.c_double(sdf.resolution), sdf.voxel_dtype))
        self._ck(self.L.stomp_engine_set_noise(self.h, _dp(_f64(scenario.noise_stddev)), _dp(_f64(scenario.noise_decay))))
        self.set_problems(self.start, self.goal)

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("stomp_b200: " + self.L.stomp_engine_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.stomp_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- distance field construction -----------------------------------------------------
    def build_sdf(self, size, origin, resolution, max_distance, boxes=(), cylinders=(), points=None, bodies=(), meshes=()):
        """boxes: (position, quaternion xyzw, dimensions); cylinders: (position, quaternion, radius, height); points: [n][3]
        collision-map points; bodies: (type, dimensions, position, quaternion, scale, padding) robot / primitive bodies in
        world poses (getVoxelsInBody); meshes: (vertices [n][3], position, quaternion, scale, padding) convex-hull mesh
        bodies (bodies::ConvexMesh)."""
        cb = (_abi.Box * max(1, len(boxes)))()
        for i, (p, q, d) in enumerate(boxes):
            cb[i].position[:], cb[i].orientation[:], cb[i].dimensions[:] = p, q, d
        cc = (_abi.Cylinder * max(1, len(cylinders)))()
        for i, (p, q, r, hgt) in enumerate(cylinders):
            cc[i].position[:], cc[i].orientation[:] = p, q
            cc[i].radius, cc[i].height = r, hgt
        bb = (_abi.Body * max(1, len(bodies)))()
        for i, (t, d, p, q, sc, pad) in enumerate(bodies):
            bb[i].type = t
            bb[i].dimensions[:] = tuple(d) + (0.0,) * (3 - len(d))
            bb[i].position[:], bb[i].orientation[:] = p, q
            bb[i].scale, bb[i].padding = sc, pad
        pts = _f64(points).reshape(-1, 3) if points is not None else None
        mm = (_abi.MeshBody * max(1, len(meshes)))()
        keep = []
        for i, (v, p, q, sc, pad) in enumerate(meshes):
            va = np.ascontiguousarray(_f64(v).reshape(-1, 3))
            keep.append(va)
            mm[i].vertices, mm[i].num_vertices = _dp(va), len(va)
            mm[i].position[:], mm[i].orientation[:] = p, q
            mm[i].scale, mm[i].padding = sc, pad
        self._ck(self.L.stomp_engine_build_sdf_meshes(self.h, (C.c_double * 3)(*size), (C.c_double * 3)(*origin), C.c_double(resolution),
                                                      C.c_double(max_distance), cb, len(boxes), cc, len(cylinders),
                                                      _dp(pts), C.c_int64(0 if pts is None else len(pts)), bb, len(bodies),
                                                      mm, len(meshes)))

    def get_sdf(self):
        dims = (C.c_int32 * 3)()
        dt = C.c_int32()
        self._ck(self.L.stomp_engine_get_sdf(self.h, dims, C.byref(dt), None, 0))
        out = np.empty(tuple(dims), dtype={_abi.VOXEL_U8_SQ: np.uint8, _abi.VOXEL_U16_SQ: np.uint16, _abi.VOXEL_F32: np.float32}[dt.value])
        self._ck(self.L.stomp_engine_get_sdf(self.h, dims, C.byref(dt), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out, dt.value

    def set_constraints(self, constraints, weight):
        """constraints: dicts {segment, orientation (x,y,z,w), tolerances (roll,pitch,yaw), weight, body_fixed}."""
        arr = (_abi.OrientationConstraint * max(1, len(constraints)))()
        for i, c in enumerate(constraints):
            arr[i].segment, arr[i].body_fixed = c["segment"], int(c.get("body_fixed", 0))
            arr[i].orientation[:] = c["orientation"]
            arr[i].absolute_roll_tolerance, arr[i].absolute_pitch_tolerance, arr[i].absolute_yaw_tolerance = c["tolerances"]
            arr[i].weight = c.get("weight", 1.0)
        self._ck(self.L.stomp_engine_set_constraints(self.h, arr, len(constraints), C.c_double(weight)))

    def set_dynamics(self, torque_cost_weight, gravity=(0.0, 0.0, -9.8)):
        """torque term of StompOptimizer::execute over the robot's inverse-dynamics chain (Robot.inertias / Robot.chain)."""
        rb = self.sc.robot
        g = _f64(gravity)
        self._ck(self.L.stomp_engine_set_dynamics(self.h, rb.c_inertias(), rb.chain[0], rb.chain[1], _dp(g),
                                                  C.c_double(torque_cost_weight)))

    def execute_constraints_satisfied(self, n):
        out = np.empty((self.B, n), dtype=np.int32)
        self._ck(self.L.stomp_engine_execute_constraints_satisfied(self.h, _ip(out), C.c_size_t(out.size)))
        return out

    # ---- policy -------------------------------------------------------------------------
    def set_problems(self, start, goal):
        s, g = _f64(start), _f64(goal)
        self._ck(self.L.stomp_engine_set_problems(self.h, _dp(s), _dp(g)))

    def set_parameters(self, theta):
        self._ck(self.L.stomp_engine_set_parameters(self.h, _dp(_f64(theta))))

    def get_parameters(self, out=None):
        if out is None:
            out = np.empty((self.B, self.D, self.N))
        self._ck(self.L.stomp_engine_get_parameters(self.h, _dp(out)))
        return out

    def update_parameters(self, updates):
        self._ck(self.L.stomp_engine_update_parameters(self.h, _dp(_f64(updates))))

    def compute_control_costs(self, parameters, noise, weight):
        p, e = _f64(parameters), _f64(noise)
        n = p.shape[1]
        out = np.empty((self.B, n, self.D, self.N))
        self._ck(self.L.stomp_engine_compute_control_costs(self.h, _dp(p), _dp(e), n, C.c_double(weight), _dp(out)))
        return out

    # ---- noise --------------------------------------------------------------------------
    def set_noise(self, noise_stddev, noise_decay):
        self._ck(self.L.stomp_engine_set_noise(self.h, _dp(_f64(noise_stddev)), _dp(_f64(noise_decay))))

    def seed(self, seed):
        self._ck(self.L.stomp_engine_seed(self.h, C.c_uint64(seed)))

    def inject_noise(self, eps):
        e = eps if (isinstance(eps, np.ndarray) and eps.dtype == np.float64 and eps.flags.c_contiguous) else _f64(eps)
        assert e.shape[0] == self.B and e.shape[2:] == (self.D, self.N)
        self._ck(self.L.stomp_engine_inject_noise(self.h, _dp(e), e.shape[1]))

    def inject_noise_async(self, eps):
        """eps must be a C-contiguous float64 array that stays alive (ideally pinned) until it has been consumed."""
        assert isinstance(eps, np.ndarray) and eps.dtype == np.float64 and eps.flags.c_contiguous
        assert eps.shape[0] == self.B and eps.shape[2:] == (self.D, self.N)
        self._ck(self.L.stomp_engine_inject_noise_async(self.h, _dp(eps), eps.shape[1]))

    def last_stats(self, cost=None, cf=None):
        cost = np.empty(self.B) if cost is None else cost
        cf = np.empty(self.B, dtype=np.int32) if cf is None else cf
        st = _abi.IterStats(_dp(cost), _ip(cf), 0, 0, None)
        self._ck(self.L.stomp_engine_last_stats(self.h, C.byref(st)))
        return cost, cf

    def request_results_async(self, theta=None, cost=None, cf=None):
        """asynchronous read-back of the last launched iteration into (pinned) numpy buffers; returns a ticket."""
        t = C.c_int32()
        self._ck(self.L.stomp_engine_request_results_async(self.h, _dp(theta), _dp(cost), _ip(cf), C.byref(t)))
        return t.value

    def wait_results(self, ticket):
        self._ck(self.L.stomp_engine_wait_results(self.h, ticket))

    def sample_noise(self, iteration, n):
        out = np.empty((self.B, n, self.D, self.N))
        self._ck(self.L.stomp_engine_sample_noise(self.h, iteration, n, _dp(out)))
        return out

    # ---- cost plugin ----------------------------------------------------------------------
    def execute(self, parameters, iteration_number=2):
        p = _f64(parameters)
        n = p.shape[1]
        costs = np.empty((self.B, n, self.N))
        cf = np.empty((self.B, n), dtype=np.int32)
        self._ck(self.L.stomp_engine_execute(self.h, _dp(p), n, iteration_number, _dp(costs), _ip(cf)))
        return costs, cf

    def execute_debug(self, parameters):
        from numpy import frombuffer
        p = _f64(parameters).reshape(self.D, self.N)
        dbg = (_abi.SphereDebug * ((self.N + 3) * self.K))()
        self._ck(self.L.stomp_engine_execute_debug(self.h, _dp(p), dbg))
        a = frombuffer(dbg, dtype=np.dtype([("voxel", np.int32, 3), ("in_collision", np.int32), ("position", np.float64, 3),
                                            ("potential", np.float64), ("vel_mag", np.float64)])).reshape(self.N + 3, self.K)
        return {k: a[k].copy() for k in a.dtype.names}

    # ---- PolicyImprovement step by step -----------------------------------------------------
    def get_rollouts(self, noise_stddev):
        out = np.empty((self.B, self.R, self.D, self.N))
        ngen = C.c_int32()
        tmp = np.empty((self.B * self.R * self.D * self.N))
        self._ck(self.L.stomp_engine_get_rollouts(self.h, _dp(_f64(noise_stddev)), _dp(tmp), C.byref(ngen)))
        g = ngen.value
        return tmp[: self.B * g * self.D * self.N].reshape(self.B, g, self.D, self.N).copy()

    def set_rollout_costs(self, costs, control_cost_weight):
        c = _f64(costs)
        totals = np.empty((self.B, self.R))
        self._ck(self.L.stomp_engine_set_rollout_costs(self.h, _dp(c), C.c_double(control_cost_weight), _dp(totals)))
        return totals

    def improve_policy(self):
        out = np.empty((self.B, self.D, self.N))
        self._ck(self.L.stomp_engine_improve_policy(self.h, _dp(out)))
        return out

    def add_extra_rollouts(self, costs):
        self._ck(self.L.stomp_engine_add_extra_rollouts(self.h, _dp(_f64(costs))))

    # ---- the hot path ---------------------------------------------------------------------
    def iterate(self, iteration_number, stats=True):
        if not stats:
            self._ck(self.L.stomp_engine_iterate(self.h, iteration_number, None))
            return None
        cost = np.empty(self.B)
        cf = np.empty(self.B, dtype=np.int32)
        st = _abi.IterStats(_dp(cost), _ip(cf), 0, 0, None)
        self._ck(self.L.stomp_engine_iterate(self.h, iteration_number, C.byref(st)))
        return cost, cf, st.num_generated_rollouts

    def run(self, first_iteration, count, stats=False):
        if not stats:
            self._ck(self.L.stomp_engine_run(self.h, first_iteration, count, None))
            return None
        cost = np.empty(self.B)
        cf = np.empty(self.B, dtype=np.int32)
        st = _abi.IterStats(_dp(cost), _ip(cf), 0, 0, None)
        self._ck(self.L.stomp_engine_run(self.h, first_iteration, count, C.byref(st)))
        return cost, cf, st.num_generated_rollouts

    def set_graph_mode(self, mode):
        """0: run() launches every kernel itself; 1 (default): steady-state iterations are replayed from a CUDA graph."""
        self._ck(self.L.stomp_engine_set_graph_mode(self.h, int(mode)))

    def optimize(self, max_iterations, max_iterations_after_collision_free):
        """StompOptimizer::optimize for the whole batch; returns a dict of per-problem statistics."""
        B = self.B
        i32 = lambda: np.full(B, -7, dtype=np.int32)
        out = dict(success=i32(), success_iteration=i32(), collision_success_iteration=i32(), last_improvement_iteration=i32(),
                   iterations=i32(), best_cost=np.zeros(B), costs=np.full((max_iterations, B), np.nan))
        st = _abi.OptimizeStats(_ip(out["success"]), _ip(out["success_iteration"]), _ip(out["collision_success_iteration"]),
                                _ip(out["last_improvement_iteration"]), _ip(out["iterations"]), _dp(out["best_cost"]), _dp(out["costs"]))
        self._ck(self.L.stomp_engine_optimize(self.h, max_iterations, max_iterations_after_collision_free, C.byref(st)))
        out["best_trajectory"] = self.get(_abi.FIELD_BEST_TRAJECTORY)
        return out

    def synchronize(self):
        self._ck(self.L.stomp_engine_synchronize(self.h))

    def timer_start(self):
        self._ck(self.L.stomp_engine_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.stomp_engine_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def set_profiling(self, enabled):
        self._ck(self.L.stomp_engine_set_profiling(self.h, int(enabled)))

    def get_profile(self, kernel_substr=""):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.L.stomp_engine_get_profile(self.h, kernel_substr.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def dump_timeline(self, path):
        """CSV of the launches recorded since set_profiling(2) (two-stream schedule kept): index, kernel, stream, begin_us, end_us."""
        self._ck(self.L.stomp_engine_dump_timeline(self.h, str(path).encode()))

    def launch_count(self):
        return int(self.L.stomp_engine_launch_count(self.h))

    def shard_buffers(self):
        mm, sm, nb = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._ck(self.L.stomp_engine_shard_buffers(self.h, C.byref(mm), C.byref(sm), C.byref(nb)))
        return mm.value, sm.value, nb.value

    def shard_ipc_handle(self):
        buf = (C.c_ubyte * 64)()
        self._ck(self.L.stomp_engine_shard_ipc_handle(self.h, buf, C.c_size_t(64)))
        return bytes(buf)

    def shard_open_peers(self, handles):
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self.L.stomp_engine_shard_open_peers(self.h, buf, len(handles)))

    def iterate_sharded_fused(self, iteration_number):
        self._ck(self.L.stomp_engine_iterate_sharded_fused(self.h, iteration_number))

    def shard_status(self):
        self._ck(self.L.stomp_engine_shard_status(self.h))

    def iterate_sharded_phase(self, iteration_number, phase):
        self._ck(self.L.stomp_engine_iterate_sharded_phase(self.h, iteration_number, phase))

    def get(self, field):
        B, R, D, N = self.B, self.R, self.D, self.N
        shapes = {
            _abi.FIELD_THETA: (B, D, N), _abi.FIELD_NOISE: (B, R, D, N), _abi.FIELD_PARAMETERS: (B, R, D, N),
            _abi.FIELD_NOISE_PROJECTED: (B, R, D, N), _abi.FIELD_STATE_COSTS: (B, R, N),
            _abi.FIELD_CONTROL_COSTS: (B, R, D, N), _abi.FIELD_CUMULATIVE_COSTS: (B, R, D, N),
            _abi.FIELD_PROBABILITIES: (B, R, D, N), _abi.FIELD_UPDATES: (B, D, N), _abi.FIELD_NOISELESS_COSTS: (B, N),
            _abi.FIELD_ROLLOUT_TOTAL_COSTS: (B, R + 1), _abi.FIELD_INV_CONTROL_COST: (N, N),
            _abi.FIELD_NOISE_CHOLESKY: (N, N), _abi.FIELD_PROJECTION: (N, N), _abi.FIELD_QUAD_COST_INV: (N, N),
            _abi.FIELD_CONTROL_COST: (N, N), _abi.FIELD_CLIPPED_PARAMETERS: (B, R, D, N),
            _abi.FIELD_BEST_TRAJECTORY: (B, D, N), _abi.FIELD_NOISELESS_TRAJECTORY: (B, D, N),
        }
        if field in (_abi.FIELD_COLLISION_FREE, _abi.FIELD_CONSTRAINTS_SATISFIED):
            out = np.empty((B, R + 1), dtype=np.int32)
        else:
            out = np.empty(shapes[field])
        self._ck(self.L.stomp_engine_get(self.h, field, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out
