"""Synthetic robots, collision spheres, baked distance fields and the BASELINE.json configs.

Nothing here is on the hot path: these are the *inputs* the reference gets from
StompRobotModel (URDF -> KDL tree, collision-sphere generation) and StompCollisionSpace
(distance field), reproduced as plain tables so that the engine and the oracle are fed
identical data.  The PR2 URDF is not part of the reference repository, so the arm below is a
stated synthetic approximation.

Reference behaviour followed:
  * collision-sphere placement every radius/2 along link -> child joint origin:
    src/stomp_robot_model.cpp:265-306
  * radii / clearance / grid box: config/pr2_both_arms_stomp_config.yaml:1-21,76-86
  * scenes: config/environment_pole.yaml, config/environment_shelf.yaml
  * parameters: config/params.yaml
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi

_I3 = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return (cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
            -sp, cp * sr, cp * cr)


@dataclass
class Robot:
    """Joint table (DFS pre-order) + collision spheres + joint limits."""
    segments: list = field(default_factory=list)   # dicts: name,parent,type,group,rot,pos,axis,fixed
    spheres: list = field(default_factory=list)    # dicts: segment,radius,clearance,pos
    limits: list = field(default_factory=list)     # (has, lo, hi) per group joint
    reference_segment: int = 0
    inertias: dict = field(default_factory=dict)   # segment -> (mass, com(3), (ixx, iyy, izz, ixy, ixz, iyz)); others massless
    chain: tuple = ()                              # (root segment, tip segment) of the inverse-dynamics chain

    @property
    def num_dimensions(self):
        return len(self.limits)

    def add_segment(self, name, parent, jtype, pos, axis=(0, 0, 1), rot=_I3, group=-1, fixed=0.0):
        a = np.asarray(axis, float)
        a = a / np.linalg.norm(a)
        self.segments.append(dict(name=name, parent=parent, type=jtype, group=group, rot=tuple(rot),
                                  pos=tuple(float(v) for v in pos), axis=tuple(a), fixed=float(fixed)))
        return len(self.segments) - 1

    def children(self, s):
        return [i for i, g in enumerate(self.segments) if g["parent"] == s]

    def add_link_spheres(self, seg, radius, clearance, extension=0.0):
        """StompRobotModel::addCollisionPointsFromLinkRadius (src/stomp_robot_model.cpp:265-306), checked against the compiled
        reference (tests/test_reference_pinning.py).  The points run from the link origin to the child's
        KDL JointOrigin().  kdl_parser gives a FIXED joint no origin (the offset lives in the segment's tip frame), so for a
        fixed child JointOrigin() is zero and the reference stacks its ceil(extension / spacing) + 1 points on the link origin;
        that is reproduced.  (With extension 0 the reference then evaluates 0 / 0; one point at the origin is emitted instead.)"""
        first_child = True
        for c in self.children(seg):
            fixed = self.segments[c]["type"] == _abi.JOINT_FIXED
            origin = np.zeros(3) if fixed else np.asarray(self.segments[c]["pos"], float)
            spacing = radius / 2.0
            distance = float(np.linalg.norm(origin)) + extension
            num_points = int(math.ceil(distance / spacing)) + 1
            for i in range(num_points):
                if not first_child and i == 0:
                    continue
                pos = origin * (i / (num_points - 1.0)) if num_points > 1 else origin * 0.0
                self.spheres.append(dict(segment=seg, radius=radius, clearance=clearance, pos=tuple(float(v) for v in pos)))
            first_child = False

    def moved_by_group(self, seg):
        """True when some planning-group joint is at or above `seg` (StompPlanningGroup::addCollisionPoint keeps only those
        points, src/stomp_robot_model.cpp:308-334)."""
        while seg >= 0:
            if self.segments[seg]["group"] >= 0:
                return True
            seg = self.segments[seg]["parent"]
        return False

    def add_even_spheres(self, seg, count, radius, clearance):
        c = self.children(seg)
        origin = np.asarray(self.segments[c[0]]["pos"], float)
        for i in range(count):
            pos = origin * (i / max(count - 1.0, 1.0))
            self.spheres.append(dict(segment=seg, radius=radius, clearance=clearance, pos=tuple(pos)))

    def add_attached_object(self, seg, shape, dimensions, position, padding=0.0, clearance=0.07):
        """StompRobotModel::generateAttachedObjectCollisionPoints (src/stomp_robot_model.cpp:377-453): one collision point per
        attached shape = its bounding sphere (geometric_shapes bodies::computeBoundingSphere, padded) in the owner link's frame.
        shape: 'sphere' (r), 'box' (x, y, z), 'cylinder' (r, length); position: shape centre in the link frame."""
        if shape == "sphere":
            radius = dimensions[0] + padding
        elif shape == "box":
            radius = math.sqrt(sum((d / 2.0 + padding) ** 2 for d in dimensions))
        elif shape == "cylinder":
            radius = math.sqrt((dimensions[0] + padding) ** 2 + (dimensions[1] / 2.0 + padding) ** 2)
        else:
            raise ValueError(shape)
        self.spheres.append(dict(segment=seg, radius=radius, clearance=clearance, pos=tuple(float(v) for v in position)))

    # ---- ctypes views -------------------------------------------------------------------
    def c_segments(self):
        arr = (_abi.Segment * len(self.segments))()
        for i, g in enumerate(self.segments):
            arr[i].parent, arr[i].joint_type, arr[i].group_index = g["parent"], g["type"], g["group"]
            arr[i].rot[:] = g["rot"]
            arr[i].pos[:] = g["pos"]
            arr[i].axis[:] = g["axis"]
            arr[i].fixed_value = g["fixed"]
        return arr

    def c_spheres(self):
        arr = (_abi.Sphere * len(self.spheres))()
        for i, s in enumerate(self.spheres):
            arr[i].segment, arr[i].radius, arr[i].clearance = s["segment"], s["radius"], s["clearance"]
            arr[i].pos[:] = s["pos"]
        return arr

    def c_inertias(self):
        """one stomp_link_inertia per segment (KDL::RigidBodyInertia of the segment in its own frame)."""
        arr = (_abi.LinkInertia * len(self.segments))()
        for i, (m, com, ic) in self.inertias.items():
            arr[i].mass = m
            arr[i].com[:] = com
            arr[i].inertia[:] = ic
        return arr

    def c_limits(self):
        arr = (_abi.JointLimit * len(self.limits))()
        for i, (has, lo, hi) in enumerate(self.limits):
            arr[i].has_limits, arr[i].min, arr[i].max = int(has), lo, hi
        return arr

    def sample_configurations(self, rng, n):
        lo = np.array([l[1] if l[0] else -math.pi for l in self.limits])
        hi = np.array([l[2] if l[0] else math.pi for l in self.limits])
        return lo + (hi - lo) * rng.random((n, len(self.limits)))


def pr2_right_arm(clearance=0.07, spheres_per_link=None, sphere_radius=0.03):
    """PR2-right-arm-like 7-DOF tree (synthetic; the URDF is not in the reference)."""
    R, F, P = _abi.JOINT_REVOLUTE, _abi.JOINT_FIXED, _abi.JOINT_PRISMATIC
    rb = Robot()
    base = rb.add_segment("base_link", -1, F, (0, 0, 0))
    torso = rb.add_segment("torso_lift_link", base, P, (-0.05, 0.0, 0.739675), (0, 0, 1), fixed=0.1)
    pan = rb.add_segment("r_shoulder_pan_link", torso, R, (0.0, -0.188, 0.0), (0, 0, 1), group=0)
    lift = rb.add_segment("r_shoulder_lift_link", pan, R, (0.1, 0.0, 0.0), (0, 1, 0), group=1)
    uroll = rb.add_segment("r_upper_arm_roll_link", lift, R, (0.0, 0.0, 0.0), (1, 0, 0), group=2)
    upper = rb.add_segment("r_upper_arm_link", uroll, F, (0.0, 0.0, 0.0))
    elbow = rb.add_segment("r_elbow_flex_link", upper, R, (0.4, 0.0, 0.0), (0, 1, 0), group=3)
    froll = rb.add_segment("r_forearm_roll_link", elbow, R, (0.0, 0.0, 0.0), (1, 0, 0), group=4)
    fore = rb.add_segment("r_forearm_link", froll, F, (0.0, 0.0, 0.0))
    wflex = rb.add_segment("r_wrist_flex_link", fore, R, (0.321, 0.0, 0.0), (0, 1, 0), group=5)
    wroll = rb.add_segment("r_wrist_roll_link", wflex, R, (0.0, 0.0, 0.0), (1, 0, 0), group=6)
    palm = rb.add_segment("r_gripper_palm_link", wroll, F, (0.0, 0.0, 0.0))
    lf = rb.add_segment("r_gripper_l_finger_link", palm, R, (0.07691, 0.01, 0.0), (0, 0, 1), fixed=0.25)
    lt = rb.add_segment("r_gripper_l_finger_tip_link", lf, R, (0.09137, 0.00495, 0.0), (0, 0, -1), fixed=0.25)
    rb.add_segment("r_gripper_l_finger_tip_frame", lt, F, (0.03, 0.0, 0.0), rot=_rpy(0.0, 0.0, 0.1))
    rf = rb.add_segment("r_gripper_r_finger_link", palm, R, (0.07691, -0.01, 0.0), (0, 0, -1), fixed=0.25)
    rt = rb.add_segment("r_gripper_r_finger_tip_link", rf, R, (0.09137, -0.00495, 0.0), (0, 0, 1), fixed=0.25)
    rb.add_segment("r_gripper_r_finger_tip_frame", rt, F, (0.03, 0.0, 0.0), rot=_rpy(0.0, 0.0, -0.1))
    rb.reference_segment = base
    rb.limits = [
        (1, -2.1353981634, 0.564601836603),
        (1, -0.3536, 1.2963),
        (1, -3.75, 0.65),
        (1, -2.1213, -0.15),
        (0, 0.0, 0.0),   # forearm roll: continuous -> no limits (src/stomp_robot_model.cpp:160-161)
        (1, -2.0, -0.1),
        (0, 0.0, 0.0),   # wrist roll: continuous
    ]
    # link inertias for the torque term: synthetic, PR2-like magnitudes (mass, centre of mass, inertia about it); the chain is
    # the one the reference hard-codes, torso_lift_link -> the end of the wrist (src/stomp_robot_model.cpp:181-183)
    rb.inertias = {
        pan: (25.8, (-0.001, -0.002, -0.27), (0.866, 0.874, 0.273, -0.061, -0.121, -0.059)),
        lift: (2.75, (0.022, -0.027, -0.031), (0.021, 0.021, 0.0198, 0.005, 0.003, 0.006)),
        uroll: (6.02, (0.21, 0.0007, -0.0003), (0.0154, 0.0775, 0.0762, 0.0002, -0.0012, -0.0001)),
        elbow: (1.9, (0.01, 0.0, -0.012), (0.0035, 0.0044, 0.0042, -0.0001, -0.0001, 0.0)),
        froll: (2.69, (0.18, -0.0003, -0.0003), (0.0146, 0.0164, 0.0146, 0.0, 0.0, 0.0)),
        wflex: (0.61, (-0.0016, 0.0, -0.0007), (0.00065, 0.0002, 0.00064, 0.0, 0.0, 0.0)),
        wroll: (0.68, (0.056, 0.0005, -0.001), (0.0012, 0.0012, 0.0004, 0.0, 0.0, 0.0)),
        palm: (0.58, (0.07, 0.0, -0.002), (0.0004, 0.0006, 0.0008, 0.0, 0.0, 0.0)),
    }
    rb.chain = (torso, palm)
    if spheres_per_link is None:
        rb.add_link_spheres(upper, 0.10, clearance)
        rb.add_link_spheres(fore, 0.065, clearance)
        rb.add_link_spheres(palm, 0.06, clearance)
        rb.add_link_spheres(lf, 0.03, clearance, 0.01)
        rb.add_link_spheres(lt, 0.03, clearance, 0.01)
        rb.add_link_spheres(rf, 0.03, clearance, 0.01)
        rb.add_link_spheres(rt, 0.03, clearance, 0.01)
    else:
        for seg in (upper, fore, palm, lf, lt, rf, rt):
            rb.add_even_spheres(seg, spheres_per_link, sphere_radius, clearance)
    return rb


def serial_chain(num_joints=30, link_length=0.1, spheres_per_link=3, radius=0.04, clearance=0.07,
                 base=(0.0, 0.0, 0.8)):
    """C5: synthetic serial chain, alternating z / y axes."""
    rb = Robot()
    parent = rb.add_segment("base_link", -1, _abi.JOINT_FIXED, (0, 0, 0))
    segs = []
    for j in range(num_joints):
        pos = base if j == 0 else (link_length, 0.0, 0.0)
        parent = rb.add_segment("joint%d" % j, parent, _abi.JOINT_REVOLUTE, pos,
                                (0, 0, 1) if j % 2 == 0 else (0, 1, 0), group=j)
        segs.append(parent)
    rb.add_segment("tool", parent, _abi.JOINT_FIXED, (link_length, 0.0, 0.0))
    for s in segs:
        rb.add_even_spheres(s, spheres_per_link, radius, clearance)
    rb.limits = [(1, -1.2, 1.2) if j % 3 else (0, 0.0, 0.0) for j in range(num_joints)]
    return rb


def random_tree(rng, num_group=5, num_extra=6, spheres=12):
    """A small random tree with rotated joint origins, prismatic and fixed joints, branching:
    used by the FK parity tests, not a benchmark config."""
    rb = Robot()
    rb.add_segment("root", -1, _abi.JOINT_FIXED, (0, 0, 0))
    kinds = [("g", j) for j in range(num_group)] + [("x", None)] * num_extra
    order = list(rng.permutation(len(kinds)))
    for k in order:
        kind, g = kinds[k]
        n = len(rb.segments)
        parent = int(rng.integers(max(0, n - 3), n))
        jt = int(rng.choice([_abi.JOINT_REVOLUTE, _abi.JOINT_REVOLUTE, _abi.JOINT_PRISMATIC])) if kind == "g" \
            else int(rng.choice([_abi.JOINT_FIXED, _abi.JOINT_REVOLUTE, _abi.JOINT_PRISMATIC]))
        rb.add_segment("s%d" % n, parent, jt, rng.uniform(-0.25, 0.25, 3), rng.normal(size=3),
                       rot=_rpy(*rng.uniform(-1.0, 1.0, 3)), group=g if kind == "g" else -1,
                       fixed=float(rng.uniform(-0.5, 0.5)))
    # renumber in DFS pre-order (children in creation order): the order kdl_parser builds a KDL::Tree in and the one
    # TreeFkSolverJointPosAxis::assignSegmentNumber (src/treefksolverjointposaxis.cpp:128-139) assigns
    order, stack = [], [0]
    while stack:
        s = stack.pop()
        order.append(s)
        stack.extend(reversed(rb.children(s)))
    new_index = {old: new for new, old in enumerate(order)}
    rb.segments = [dict(rb.segments[old], parent=new_index.get(rb.segments[old]["parent"], -1)) for old in order]
    # reference frame: a segment that no group joint moves (first segment whose ancestry is static)
    rb.reference_segment = 0
    for _ in range(spheres):
        rb.spheres.append(dict(segment=int(rng.integers(0, len(rb.segments))), radius=float(rng.uniform(0.02, 0.1)),
                               clearance=0.07, pos=tuple(rng.uniform(-0.2, 0.2, 3))))
    rb.limits = [(int(rng.integers(0, 2)), -0.8, 0.9) for _ in range(num_group)]
    return rb


def fill_in_min_jerk(start, goal, num_free_points, discretization):
    """StompTrajectory::fillInMinJerk (src/stomp_trajectory.cpp:179-223): quintic with zero start / end velocity and
    acceleration between the fixed points before and after the free block.  Returns [D][num_free_points]."""
    start, goal = np.asarray(start, float), np.asarray(goal, float)
    T1 = (num_free_points + 1) * discretization
    T = [T1 ** k for k in range(6)]
    c3 = (-20 * start + 20 * goal) / (2 * T[3])
    c4 = (30 * start - 30 * goal) / (2 * T[4])
    c5 = (-12 * start + 12 * goal) / (2 * T[5])
    t = (np.arange(1, num_free_points + 1) * discretization)[None, :]
    return start[:, None] + c3[:, None] * t ** 3 + c4[:, None] * t ** 4 + c5[:, None] * t ** 5


# ---- distance fields ---------------------------------------------------------------------

@dataclass
class DistanceField:
    voxels: np.ndarray          # [nx][ny][nz], C-contiguous
    origin: tuple
    resolution: float
    voxel_dtype: int

    @property
    def dims(self):
        return self.voxels.shape


def bake_distance_field(size, origin, resolution, boxes=(), cylinders=(), max_distance=0.17,
                        voxel_dtype=_abi.VOXEL_U8_SQ):
    """Squared-cell-distance grid like distance_field::PropagationDistanceField:
    num_cells = int(size/res); voxel = min(cell distance^2, ceil(max_distance/res)^2), exact EDT."""
    from scipy import ndimage

    n = [int(size[i] / resolution) for i in range(3)]
    ax = [origin[i] + resolution * np.arange(n[i]) for i in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij", sparse=True)
    occ = np.zeros(n, dtype=bool)
    h = 0.5 * resolution
    for (c, d) in boxes:        # centre, dimensions
        occ |= ((np.abs(X - c[0]) <= d[0] / 2 + h) & (np.abs(Y - c[1]) <= d[1] / 2 + h) &
                (np.abs(Z - c[2]) <= d[2] / 2 + h))
    for (c, r, hgt) in cylinders:  # centre, radius, height (axis z)
        occ |= (((X - c[0]) ** 2 + (Y - c[1]) ** 2 <= (r + h) ** 2) & (np.abs(Z - c[2]) <= hgt / 2 + h))
    cap = int(math.ceil(max_distance / resolution))
    if occ.any():
        d = ndimage.distance_transform_edt(~occ)
        d2 = np.minimum(np.rint(d * d), cap * cap).astype(np.int64)
    else:
        d2 = np.full(n, cap * cap, dtype=np.int64)
    if voxel_dtype == _abi.VOXEL_U8_SQ:
        assert cap * cap < 256
        vox = d2.astype(np.uint8)
    elif voxel_dtype == _abi.VOXEL_U16_SQ:
        vox = d2.astype(np.uint16)
    else:
        vox = (np.sqrt(d2.astype(np.float64)) * resolution).astype(np.float32)
    return DistanceField(np.ascontiguousarray(vox), tuple(origin), float(resolution), voxel_dtype)


POLE = dict(cylinders=[((0.62, -0.62, 0.6), 0.1, 1.2)])   # config/environment_pole.yaml:4-10


def _shelf_boxes():
    """config/environment_shelf.yaml:4-64: six shelf boards and four uprights."""
    boxes = []
    for z in (0.015, 0.329, 0.643, 0.957, 1.271, 1.585):
        boxes.append(((0.8, -0.1, z), (0.4, 1.2, 0.03)))
    for y in (-0.685, -0.295, 0.095, 0.485):
        boxes.append(((0.8, y, 0.8), (0.4, 0.03, 1.6)))
    return boxes


SHELF = dict(boxes=_shelf_boxes())


def random_clutter(rng, count, lo, hi, min_size=0.03, max_size=0.15, keep_out=None):
    boxes = []
    while len(boxes) < count:
        c = rng.uniform(lo, hi)
        d = rng.uniform(min_size, max_size, 3)
        if keep_out is not None and np.linalg.norm(c - keep_out[0]) < keep_out[1]:
            continue
        boxes.append((tuple(c), tuple(d)))
    return boxes


# ---- BASELINE.json configurations ---------------------------------------------------------

@dataclass
class Scenario:
    name: str
    robot: Robot
    sdf: DistanceField
    num_time_steps: int
    num_rollouts: int
    num_reused_rollouts: int
    num_problems: int
    start: np.ndarray            # [B][D]
    goal: np.ndarray             # [B][D]
    noise_stddev: np.ndarray
    noise_decay: np.ndarray
    movement_duration: float = 5.0
    discretization: float = 0.05
    derivative_costs: tuple = (0.0, 1.0, 0.0)
    ridge_factor: float = 0.0
    smoothness_cost_weight: float = 1e-6
    obstacle_cost_weight: float = 1.0
    use_cumulative_costs: int = 0
    sdf_mode: int = _abi.SDF_NEAREST   # SDF_TRILINEAR: the engine's interpolated extension (not a reference mode)

    def desc(self, dtype=_abi.F64, device=0, keep_intermediates=0, num_problems=None,
             shard_rank=0, shard_world=1):
        d = _abi.EngineDesc()
        d.num_dimensions = self.robot.num_dimensions
        d.num_time_steps = self.num_time_steps
        d.num_rollouts = self.num_rollouts
        d.num_reused_rollouts = self.num_reused_rollouts
        d.num_problems = self.num_problems if num_problems is None else num_problems
        d.dtype = dtype
        d.use_cumulative_costs = self.use_cumulative_costs
        d.sdf_mode = self.sdf_mode
        d.device = device
        d.rollout_shard_rank = shard_rank
        d.rollout_shard_world = shard_world
        d.keep_intermediates = keep_intermediates
        d.movement_duration = self.movement_duration
        d.discretization = self.discretization
        d.derivative_costs[:] = self.derivative_costs
        d.ridge_factor = self.ridge_factor
        d.smoothness_cost_weight = self.smoothness_cost_weight
        d.obstacle_cost_weight = self.obstacle_cost_weight
        return d


_C1_GRID = dict(size=(2.0, 3.0, 2.2), origin=(-0.5, -1.5, -0.3), resolution=0.015)  # pr2_both_arms_stomp_config.yaml:76-86

_sdf_cache = {}


def _cached_sdf(key, **kw):
    if key not in _sdf_cache:
        _sdf_cache[key] = bake_distance_field(**kw)
    return _sdf_cache[key]


def make_scenario(name, num_problems=None, num_time_steps=None, num_rollouts=None, seed=7, scene="shelf+pole",
                  use_cumulative_costs=0):
    """name: 'C1' .. 'C5' (BASELINE.json configs[0..4]) or 'tiny' (unit-test size)."""
    rng = np.random.default_rng(seed)
    if name in ("C1", "C2", "C3"):
        rb = pr2_right_arm()
        obstacles = dict(boxes=SHELF["boxes"] if "shelf" in scene else [], cylinders=POLE["cylinders"] if "pole" in scene else [])
        sdf = _cached_sdf(("c1", scene), **_C1_GRID, **obstacles, max_distance=0.17)
        N = num_time_steps or 100
        if name == "C1":
            B, R, Rre = num_problems or 1, num_rollouts or 10, 5
        elif name == "C2":
            B, R, Rre = num_problems or 1024, num_rollouts or 10, 5
        else:
            B, R, Rre = num_problems or 1, num_rollouts or 65536, 0
    elif name == "C4":
        rb = pr2_right_arm(spheres_per_link=60, sphere_radius=0.03)
        boxes = random_clutter(np.random.default_rng(11), 512, np.array([-0.2, -1.2, 0.0]), np.array([2.3, 1.3, 2.2]),
                               keep_out=(np.array([-0.05, -0.188, 0.84]), 0.35))
        sdf = _cached_sdf(("c4",), size=(2.56, 2.56, 2.56), origin=(-0.25, -1.28, -0.2), resolution=0.01, boxes=boxes,
                          max_distance=0.10)
        N = num_time_steps or 200
        B, R, Rre = num_problems or 4096, num_rollouts or 10, 5
    elif name == "C5":
        rb = serial_chain(30)
        sdf = _cached_sdf(("c1", scene), **_C1_GRID, boxes=SHELF["boxes"], cylinders=POLE["cylinders"], max_distance=0.17)
        N = num_time_steps or 300
        B, R, Rre = num_problems or 1, num_rollouts or 512, 0
    elif name == "tiny":
        rb = pr2_right_arm()
        sdf = _cached_sdf(("tiny",), size=(1.6, 1.6, 1.6), origin=(-0.3, -1.0, 0.0), resolution=0.04,
                          boxes=[((0.7, -0.3, 0.7), (0.3, 0.5, 0.06))], cylinders=[((0.5, -0.6, 0.8), 0.06, 1.0)],
                          max_distance=0.17)
        N = num_time_steps or 20
        B, R, Rre = num_problems or 2, num_rollouts or 6, 3
    else:
        raise ValueError(name)
    D = rb.num_dimensions
    start = rb.sample_configurations(rng, B)
    goal = rb.sample_configurations(rng, B)
    return Scenario(name=name, robot=rb, sdf=sdf, num_time_steps=N, num_rollouts=R, num_reused_rollouts=Rre,
                    num_problems=B, start=start, goal=goal, noise_stddev=np.full(D, 2.0), noise_decay=np.full(D, 0.999),
                    use_cumulative_costs=use_cumulative_costs)
