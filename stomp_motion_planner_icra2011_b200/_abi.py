"""ctypes mirror of include/stomp_b200.h (structs, enums).  Plain data only."""
import ctypes as C

DIFF_RULE_LENGTH = 7
NUM_DIFF_RULES = 3

F64, F32 = 0, 1
SDF_NEAREST, SDF_TRILINEAR = 0, 1
VOXEL_F32, VOXEL_U8_SQ, VOXEL_U16_SQ = 0, 1, 2
JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2

(FIELD_THETA, FIELD_NOISE, FIELD_PARAMETERS, FIELD_NOISE_PROJECTED, FIELD_STATE_COSTS, FIELD_CONTROL_COSTS,
 FIELD_CUMULATIVE_COSTS, FIELD_PROBABILITIES, FIELD_UPDATES, FIELD_NOISELESS_COSTS, FIELD_COLLISION_FREE,
 FIELD_ROLLOUT_TOTAL_COSTS, FIELD_INV_CONTROL_COST, FIELD_NOISE_CHOLESKY, FIELD_PROJECTION, FIELD_QUAD_COST_INV,
 FIELD_CONTROL_COST, FIELD_CLIPPED_PARAMETERS, FIELD_BEST_TRAJECTORY, FIELD_NOISELESS_TRAJECTORY,
 FIELD_CONSTRAINTS_SATISFIED) = range(21)


class EngineDesc(C.Structure):
    _fields_ = [
        ("num_dimensions", C.c_int32),
        ("num_time_steps", C.c_int32),
        ("num_rollouts", C.c_int32),
        ("num_reused_rollouts", C.c_int32),
        ("num_problems", C.c_int32),
        ("dtype", C.c_int32),
        ("use_cumulative_costs", C.c_int32),
        ("sdf_mode", C.c_int32),
        ("device", C.c_int32),
        ("rollout_shard_rank", C.c_int32),
        ("rollout_shard_world", C.c_int32),
        ("keep_intermediates", C.c_int32),
        ("movement_duration", C.c_double),
        ("discretization", C.c_double),
        ("derivative_costs", C.c_double * NUM_DIFF_RULES),
        ("ridge_factor", C.c_double),
        ("smoothness_cost_weight", C.c_double),
        ("obstacle_cost_weight", C.c_double),
    ]


class Segment(C.Structure):
    _fields_ = [
        ("parent", C.c_int32),
        ("joint_type", C.c_int32),
        ("group_index", C.c_int32),
        ("reserved0", C.c_int32),
        ("rot", C.c_double * 9),
        ("pos", C.c_double * 3),
        ("axis", C.c_double * 3),
        ("fixed_value", C.c_double),
    ]


class Sphere(C.Structure):
    _fields_ = [
        ("segment", C.c_int32),
        ("reserved0", C.c_int32),
        ("radius", C.c_double),
        ("clearance", C.c_double),
        ("pos", C.c_double * 3),
    ]


class JointLimit(C.Structure):
    _fields_ = [
        ("has_limits", C.c_int32),
        ("reserved0", C.c_int32),
        ("min", C.c_double),
        ("max", C.c_double),
    ]


class IterStats(C.Structure):
    _fields_ = [
        ("noiseless_cost", C.POINTER(C.c_double)),
        ("noiseless_collision_free", C.POINTER(C.c_int32)),
        ("num_generated_rollouts", C.c_int32),
        ("reserved0", C.c_int32),
        ("noiseless_constraints_satisfied", C.POINTER(C.c_int32)),
    ]


class OrientationConstraint(C.Structure):
    _fields_ = [
        ("segment", C.c_int32),
        ("body_fixed", C.c_int32),
        ("orientation", C.c_double * 4),
        ("absolute_roll_tolerance", C.c_double),
        ("absolute_pitch_tolerance", C.c_double),
        ("absolute_yaw_tolerance", C.c_double),
        ("weight", C.c_double),
    ]


class LinkInertia(C.Structure):
    _fields_ = [("mass", C.c_double), ("com", C.c_double * 3), ("inertia", C.c_double * 6)]


class SphereDebug(C.Structure):
    _fields_ = [
        ("voxel", C.c_int32 * 3),
        ("in_collision", C.c_int32),
        ("position", C.c_double * 3),
        ("potential", C.c_double),
        ("vel_mag", C.c_double),
    ]


class OptimizeStats(C.Structure):
    _fields_ = [
        ("success", C.POINTER(C.c_int32)),
        ("success_iteration", C.POINTER(C.c_int32)),
        ("collision_success_iteration", C.POINTER(C.c_int32)),
        ("last_improvement_iteration", C.POINTER(C.c_int32)),
        ("iterations", C.POINTER(C.c_int32)),
        ("best_cost", C.POINTER(C.c_double)),
        ("costs", C.POINTER(C.c_double)),
    ]


class Box(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("orientation", C.c_double * 4), ("dimensions", C.c_double * 3)]


class Cylinder(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("orientation", C.c_double * 4), ("radius", C.c_double), ("height", C.c_double)]


BODY_SPHERE, BODY_BOX, BODY_CYLINDER = 0, 1, 2


class MeshBody(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_double)), ("num_vertices", C.c_int32), ("reserved", C.c_int32),
                ("position", C.c_double * 3), ("orientation", C.c_double * 4), ("scale", C.c_double), ("padding", C.c_double)]


class Body(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("dimensions", C.c_double * 3), ("position", C.c_double * 3),
                ("orientation", C.c_double * 4), ("scale", C.c_double), ("padding", C.c_double)]
