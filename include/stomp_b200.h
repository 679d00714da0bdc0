/*
 * stomp_b200.h — C ABI of the B200-native STOMP rollout engine.
 *
 * This is the drop-in boundary for the per-iteration rollout loop of
 * kalakris/stomp_motion_planner_icra2011.  Every entry point names the
 * reference interface it replaces (paths relative to
 * stomp_motion_planner/ in the reference tree).  Plain pointers and sizes
 * only; the caller owns every host buffer, the engine owns device memory.
 *
 * Conventions
 *   D  = joints in the planning group   (num_dimensions)
 *   N  = free trajectory points         (num_time_steps; params == timesteps)
 *   R  = rollouts per iteration         (num_rollouts)
 *   K  = collision spheres of the group
 *   B  = independent planning problems held by one engine (reference: 1)
 *   All host arrays are C-contiguous fp64 with the LAST index fastest:
 *   theta[B][D][N], rollouts[B][R][D][N], costs[B][R][N].
 *   Every function returns 0 on success, non-zero on failure; the message is
 *   available from stomp_engine_last_error().  (Reference convention: every
 *   method returns bool and logs with ROS_ERROR, include/.../assert.h:43-56.)
 *   A handle is not thread-safe: one handle = one device = one CUDA stream.
 */
#ifndef STOMP_B200_H_
#define STOMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STOMP_DIFF_RULE_LENGTH 7 /* include/stomp_motion_planner/stomp_utils.h:49 */
#define STOMP_NUM_DIFF_RULES 3   /* include/stomp_motion_planner/stomp_utils.h:50 */

enum stomp_dtype { STOMP_F64 = 0, STOMP_F32 = 1 };
/* NEAREST: the reference's lookup (distance of the nearest cell, 0 on / outside the grid's outermost layer): the parity mode.
 * TRILINEAR: engine extension, trilinear interpolation of the eight surrounding cell distances (each corner with the same
 * boundary rule): a continuous field, no parity claim against the reference. */
enum stomp_sdf_mode { STOMP_SDF_NEAREST = 0, STOMP_SDF_TRILINEAR = 1 };
enum stomp_voxel_dtype {
  STOMP_VOXEL_F32 = 0,   /* distance in metres */
  STOMP_VOXEL_U8_SQ = 1, /* squared cell distance, distance = sqrt(v)*resolution */
  STOMP_VOXEL_U16_SQ = 2
};
enum stomp_joint_type { STOMP_JOINT_FIXED = 0, STOMP_JOINT_REVOLUTE = 1, STOMP_JOINT_PRISMATIC = 2 };
enum stomp_noise_mode {
  STOMP_NOISE_PHILOX = 0,  /* engine RNG: eps = sigma * C^-T z (R = C C^T, banded), z ~ Philox4x32-10 + fp32 Box-Muller */
  STOMP_NOISE_INJECTED = 1 /* host-injection mode: eps supplied by stomp_engine_inject_noise */
};

/* Scalar configuration.  Replaces the ROS parameters read by
 * PolicyImprovementLoop::readParameters (src/policy_improvement_loop.cpp:112-123),
 * StompParameters::initFromNodeHandle (src/stomp_parameters.cpp:50-76) and the
 * arguments of CovariantTrajectoryPolicy::initialize
 * (src/covariant_trajectory_policy.cpp:70-90). */
typedef struct stomp_engine_desc {
  int32_t num_dimensions;
  int32_t num_time_steps;
  int32_t num_rollouts;
  int32_t num_reused_rollouts;
  int32_t num_problems;
  int32_t dtype;                /* stomp_dtype: arithmetic type of the device path */
  int32_t use_cumulative_costs; /* src/policy_improvement_loop.cpp:121 */
  int32_t sdf_mode;             /* stomp_sdf_mode; NEAREST is the parity mode */
  int32_t device;               /* CUDA device ordinal */
  int32_t rollout_shard_rank;   /* rollout sharding of ONE problem over GPUs (config C3) */
  int32_t rollout_shard_world;  /* 1 = no sharding */
  int32_t keep_intermediates;  /* 1: also store the per-rollout parity taps (noise_projected, cumulative
                                  costs, probabilities) that the fused update never needs in HBM */
  double movement_duration;     /* policy dt = duration/(N+1), covariant_trajectory_policy.cpp:152 */
  double discretization;        /* trajectory discretization: FD velocity (stomp_optimizer.cpp:620), StompCost */
  double derivative_costs[STOMP_NUM_DIFF_RULES]; /* vel, acc, jerk */
  double ridge_factor;
  double smoothness_cost_weight; /* == control cost weight, stomp_optimizer.cpp:1178-1182 */
  double obstacle_cost_weight;
} stomp_engine_desc;

/* One KDL-tree segment in DFS pre-order (parent index < own index).
 * Replaces the kdl_parser/KDL::Tree consumed by
 * KDL::TreeFkSolverJointPosAxisPartial (src/treefksolverjointposaxis_partial.cpp:76-178).
 * segment.pose(q) = Frame(Rot(axis, q) * rot, pos)          (revolute)
 *                   Frame(rot, pos + q * axis)              (prismatic)
 *                   Frame(rot, pos)                         (fixed) */
typedef struct stomp_segment {
  int32_t parent;      /* -1 for the root */
  int32_t joint_type;  /* stomp_joint_type */
  int32_t group_index; /* [0,D) if the joint belongs to the planning group, else -1 */
  int32_t reserved0;
  double rot[9];       /* parent->joint rotation, row-major */
  double pos[3];       /* parent->joint translation (== KDL JointOrigin) */
  double axis[3];      /* joint axis in the parent segment frame (== KDL JointAxis) */
  double fixed_value;  /* joint value when the joint is not in the group (robot start state) */
} stomp_segment;

/* Replaces StompCollisionPoint (include/stomp_motion_planner/stomp_collision_point.h:76-82). */
typedef struct stomp_sphere {
  int32_t segment;
  int32_t reserved0;
  double radius;
  double clearance;
  double pos[3]; /* centre in the segment frame */
} stomp_sphere;

/* Replaces StompJoint limits (include/stomp_motion_planner/stomp_robot_model.h:78-90). */
typedef struct stomp_joint_limit {
  int32_t has_limits;
  int32_t reserved0;
  double min;
  double max;
} stomp_joint_limit;

/* Replaces OrientationConstraintEvaluator (src/constraint_evaluator.cpp:50-114): per free trajectory point
 *   cost = weight * (rw*|roll| + pw*|pitch| + yw*|yaw|)   of   R_segment * R_nominal^-1  (header frame)
 *                                                         or   R_nominal^-1 * R_segment  (body fixed),
 * roll/pitch/yaw as bullet's btMatrix3x3::getRPY returns them, rw/pw/yw = 0 when the matching tolerance is >= pi, and
 * the point "satisfies" the constraint when every angle is within its tolerance.  The per-timestep state cost gets
 * constraint_cost_weight * sum of the constraint costs added (src/stomp_optimizer.cpp:1107-1151). */
typedef struct stomp_orientation_constraint {
  int32_t segment;     /* segment whose frame is constrained (frame_number_) */
  int32_t body_fixed;  /* 0: HEADER_FRAME, 1: body-fixed */
  double orientation[4]; /* nominal orientation, quaternion x, y, z, w */
  double absolute_roll_tolerance, absolute_pitch_tolerance, absolute_yaw_tolerance;
  double weight;
} stomp_orientation_constraint;

/* Per-iteration result; mirrors what StompOptimizer::optimize reads after
 * runSingleIteration (src/stomp_optimizer.cpp:301-339): last_trajectory_cost_,
 * last_trajectory_collision_free_.  Arrays are [B], caller-allocated, may be NULL. */
typedef struct stomp_iter_stats {
  double* noiseless_cost;          /* sum_t costs of the noise-less rollout */
  int32_t* noiseless_collision_free;
  int32_t num_generated_rollouts;  /* R on the first iteration, R - R_reuse afterwards */
  int32_t reserved0;
  int32_t* noiseless_constraints_satisfied; /* last_trajectory_constraints_satisfied_, [B], may be NULL */
} stomp_iter_stats;

/* Parity taps / state getters, selector for stomp_engine_get. */
enum stomp_field {
  STOMP_FIELD_THETA = 0,          /* [B][D][N]     policy parameters (free block of parameters_all_) */
  STOMP_FIELD_NOISE = 1,          /* [B][R][D][N]  Rollout::noise_ */
  STOMP_FIELD_PARAMETERS = 2,     /* [B][R][D][N]  Rollout::parameters_ */
  STOMP_FIELD_NOISE_PROJECTED = 3,/* [B][R][D][N]  Rollout::noise_projected_ */
  STOMP_FIELD_STATE_COSTS = 4,    /* [B][R][N]     Rollout::state_costs_ */
  STOMP_FIELD_CONTROL_COSTS = 5,  /* [B][R][D][N]  Rollout::control_costs_ */
  STOMP_FIELD_CUMULATIVE_COSTS = 6,/*[B][R][D][N]  Rollout::cumulative_costs_ */
  STOMP_FIELD_PROBABILITIES = 7,  /* [B][R][D][N]  Rollout::probabilities_ */
  STOMP_FIELD_UPDATES = 8,        /* [B][D][N]     row 0 of parameter_updates_ */
  STOMP_FIELD_NOISELESS_COSTS = 9,/* [B][N]        tmp_rollout_cost_ of the noise-less rollout */
  STOMP_FIELD_COLLISION_FREE = 10,/* [B][R+1] int32 (slot R = noise-less rollout) */
  STOMP_FIELD_ROLLOUT_TOTAL_COSTS = 11, /* [B][R+1] Rollout::getCost() (slot R = extra rollout) */
  STOMP_FIELD_INV_CONTROL_COST = 12, /* [N][N] R^-1 */
  STOMP_FIELD_NOISE_CHOLESKY = 13,   /* [N][N] lower chol(R^-1) */
  STOMP_FIELD_PROJECTION = 14,       /* [N][N] M */
  STOMP_FIELD_QUAD_COST_INV = 15,    /* [N][N] StompCost::quad_cost_inv_ (scaled) */
  STOMP_FIELD_CONTROL_COST = 16,     /* [N][N] R */
  STOMP_FIELD_CLIPPED_PARAMETERS = 17, /* [B][R][D][N] trajectory after handleJointLimits (what FK sees) */
  STOMP_FIELD_BEST_TRAJECTORY = 18,    /* [B][D][N] best_group_trajectory_ kept by stomp_engine_optimize */
  STOMP_FIELD_NOISELESS_TRAJECTORY = 19, /* [B][D][N] group trajectory of the last noise-less rollout (after joint limits) */
  STOMP_FIELD_CONSTRAINTS_SATISFIED = 20 /* [B][R+1] int32 last_trajectory_constraints_satisfied_ (slot R = noise-less rollout) */
};

/* Per-sphere debug record of one rollout (parity tap for the integer work). */
typedef struct stomp_sphere_debug {
  int32_t voxel[3];     /* grid cell, PropagationDistanceField::worldToGrid */
  int32_t in_collision; /* point_is_in_collision_ */
  double position[3];   /* collision_point_pos_ */
  double potential;     /* collision_point_potential_ */
  double vel_mag;       /* collision_point_vel_mag_ */
} stomp_sphere_debug;

/* ---- lifetime ------------------------------------------------------------------ */
int stomp_engine_create(const stomp_engine_desc* desc, void** out_engine);
int stomp_engine_destroy(void* engine);
const char* stomp_engine_last_error(void);
/* ABI version and a constant that proves the CUDA build is the one loaded. */
int stomp_engine_abi_version(void);
const char* stomp_engine_build_info(void);

/* ---- scene / robot (outputs of StompRobotModel and StompCollisionSpace) ---------- */
/* Replaces StompRobotModel::StompPlanningGroup {fk_solver_, collision_points_, stomp_joints_}
 * (include/stomp_motion_planner/stomp_robot_model.h:95-119). */
int stomp_engine_set_robot(void* engine, const stomp_segment* segments, int32_t num_segments,
                           int32_t reference_segment, const stomp_sphere* spheres, int32_t num_spheres,
                           const stomp_joint_limit* limits /* [D] */);
/* Replaces the distance_field::PropagationDistanceField owned by StompCollisionSpace
 * (src/stomp_collision_space.cpp:83).  voxels is x-major [nx][ny][nz].  For the *_SQ voxel
 * types distance = sqrt(v) * resolution. */
int stomp_engine_set_sdf(void* engine, const void* voxels, int32_t nx, int32_t ny, int32_t nz,
                         const double origin[3], double resolution, int32_t voxel_dtype);
/* Rebuilds the distance field on the device from collision objects, like StompCollisionSpace::setStartState does per
 * planning request (src/stomp_collision_space.cpp:154-197): every box / cylinder is sampled on a `resolution` lattice
 * exactly as addCollisionObjectsToPoints does (src/stomp_collision_space.cpp:238-293: lattice from the low corner up to
 * dimension + resolution, point = pose * (position - lattice point)), the points are binned with the field's
 * worldToGrid rule, and the squared cell distance to the nearest occupied cell, capped at ceil(max_distance /
 * resolution)^2, is computed for every voxel — the EXACT Euclidean distance transform.  The un-vendored ROS
 * PropagationDistanceField approximates the same quantity by propagation: restated in the oracle, it differs from the exact
 * transform in ~2e-5 of the cells of the benchmark scene, by at most 2 squared cells (tests/test_oracle_kats.py).
 * num_cells = int(size / resolution).  The result replaces the
 * field set by stomp_engine_set_sdf (u8 voxels when the cap fits, else u16). */
typedef struct stomp_box {
  double position[3];
  double orientation[4]; /* quaternion x, y, z, w */
  double dimensions[3];
} stomp_box;
typedef struct stomp_cylinder {
  double position[3];
  double orientation[4];
  double radius;
  double height;
} stomp_cylinder;
int stomp_engine_build_sdf(void* engine, const double size[3], const double origin[3], double resolution, double max_distance,
                           const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders);
/* The same with collision-map points added: every point[num_points][3] (reference frame, metres) occupies the cell it falls
 * into — the "points" namespace of the environment model (src/stomp_collision_space.cpp:205-213), e.g. a sensor's voxel map. */
int stomp_engine_build_sdf_points(void* engine, const double size[3], const double origin[3], double resolution, double max_distance,
                                  const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                  const double* points, int64_t num_points);
/* The same with robot bodies added: StompCollisionSpace::setStartState also voxelises the robot's own links that are not in the
 * planning group (torso, head, the other arm ...) at the start state, and primitive collision objects handed over as bodies
 * (src/stomp_collision_space.cpp:167-188, addAllBodiesButExcludeLinksToPoints :567-588, getVoxelsInBody :590-650).
 * getVoxelsInBody walks the lattice  centre + k * resolution,  |k| <= int(bounding radius / resolution)  per axis around the
 * body's bounding-sphere centre and keeps the points from which a +z ray crosses the body's surface an odd number of times;
 * for the convex primitives here that is "strictly inside the scaled + padded shape" (geometric_shapes' bodies::Sphere / Box /
 * Cylinder are not vendored: their scale-then-pad convention and bounding spheres are restated).  Every
 * kept point occupies the cell int(round((p - origin) / resolution)) like every other point.  `bodies` carry WORLD poses (the
 * caller evaluates the links' transforms at the start state; the facade's StompCollisionSpace::setStartState does). */
enum stomp_body_type { STOMP_BODY_SPHERE = 0, STOMP_BODY_BOX = 1, STOMP_BODY_CYLINDER = 2 };
typedef struct stomp_body {
  int32_t type;
  int32_t reserved;
  double dimensions[3];  /* sphere: radius; box: x, y, z extents; cylinder: radius, length (axis z) */
  double position[3];
  double orientation[4]; /* quaternion x, y, z, w */
  double scale;          /* bodies::Body::setScale, 1.0 = unscaled */
  double padding;        /* bodies::Body::setPadding, metres */
} stomp_body;
int stomp_engine_build_sdf_bodies(void* engine, const double size[3], const double origin[3], double resolution, double max_distance,
                                  const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                  const double* points, int64_t num_points, const stomp_body* bodies, int32_t num_bodies);
/* The same with MESH bodies added (shapes::MESH collision objects, src/stomp_collision_space.cpp:217-223, and robot links whose
 * collision geometry is a mesh).  bodies::createBodyFromShape turns a mesh into a bodies::ConvexMesh = the convex hull of its
 * vertices (geometric_shapes, not vendored; restated from its published source): the hull's vertices give the
 * mesh centre (their mean) and the bounding radius (largest distance from it); scale and padding move every hull vertex along
 * its ray from the centre, v' = centre + (v - centre) * (scale + padding / |v - centre|); the bounding sphere is
 * (pose * centre, radius * scale + padding).  getVoxelsInBody then keeps the lattice points of that sphere's box from which a
 * +z ray crosses the hull's triangles an odd number of times (:625-646).  The engine computes the hull on the host, the ray
 * parity on the device (k_sdf_mark_meshes; shared edges are assigned to exactly one triangle, so a ray through an edge or a
 * vertex still counts once).  `vertices` are in the body frame; triangles of the input mesh are not needed. */
typedef struct stomp_mesh_body {
  const double* vertices; /* [num_vertices][3] */
  int32_t num_vertices;
  int32_t reserved;
  double position[3];
  double orientation[4];  /* quaternion x, y, z, w */
  double scale;           /* 1.0 = unscaled */
  double padding;         /* metres */
} stomp_mesh_body;
int stomp_engine_build_sdf_meshes(void* engine, const double size[3], const double origin[3], double resolution, double max_distance,
                                  const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                  const double* points, int64_t num_points, const stomp_body* bodies, int32_t num_bodies,
                                  const stomp_mesh_body* meshes, int32_t num_meshes);
/* Copies the current voxel grid out (parity tap): dims[3], voxel dtype, and up to `bytes` of voxels (may be NULL). */
int stomp_engine_get_sdf(void* engine, int32_t dims[3], int32_t* voxel_dtype, void* voxels, size_t bytes);
/* Inverse-dynamics (torque) cost term of StompOptimizer::execute (src/stomp_optimizer.cpp:1117-1142, getTorques :1006-1061):
 *   costs(t) += torque_cost_weight * sum_j |tau_j(t)|,   tau = RNE(q(t), qd(t), qdd(t)) with zero external wrenches,
 * q = the joint-limit-projected group trajectory, qd / qdd = its 7-tap finite differences (DIFF_RULES rows 0 and 1 over the
 * padded trajectory, include/stomp_motion_planner/stomp_trajectory.h:286-310), RNE = KDL::ChainIdSolver_RNE over the
 * KDL::Chain of the planning group (src/stomp_robot_model.cpp:181-185) with gravity (0, 0, -9.8) in the chain root's frame.
 * The chain is the path chain_root_segment -> chain_tip_segment of the segment table (root excluded, like
 * KDL::Tree::getChain); its movable joints must be exactly the group joints 0..D-1 in order (the reference indexes the chain
 * solver's joint arrays by group joint).  One stomp_link_inertia per segment of the table = the KDL::RigidBodyInertia
 * kdl_parser attaches to that segment: mass, centre of mass and rotational inertia about the centre of mass, all in the
 * segment's own (link) frame.  The term is evaluated only when torque_cost_weight > 1e-9 (the reference's test); weight 0
 * (every shipped configuration, src/stomp_parameters.cpp:56) removes it. */
typedef struct stomp_link_inertia {
  double mass;
  double com[3];
  double inertia[6]; /* ixx, iyy, izz, ixy, ixz, iyz */
} stomp_link_inertia;
int stomp_engine_set_dynamics(void* engine, const stomp_link_inertia* inertia /* [num_segments] */, int32_t chain_root_segment,
                              int32_t chain_tip_segment, const double gravity[3], double torque_cost_weight);
/* Orientation path constraints of the planning request (StompOptimizer::initialize, src/stomp_optimizer.cpp:196-201) and
 * constraint_cost_weight (config/params.yaml:12).  n == 0 removes them.  Must be called after stomp_engine_set_robot. */
int stomp_engine_set_constraints(void* engine, const stomp_orientation_constraint* constraints, int32_t n,
                                 double constraint_cost_weight);
/* noise_stddev / noise_decay of PolicyImprovementLoop (src/policy_improvement_loop.cpp:118-119,155-160). */
int stomp_engine_set_noise(void* engine, const double* noise_stddev /* [D] */, const double* noise_decay /* [D] */);

/* ---- policy (CovariantTrajectoryPolicy) --------------------------------------- */
/* setToMinControlCost(start, goal) for every problem (src/covariant_trajectory_policy.cpp:102-148);
 * also resets the rollout-reuse state like PolicyImprovement::initialize (src/policy_improvement.cpp:64-94). */
int stomp_engine_set_problems(void* engine, const double* start /* [B][D] */, const double* goal /* [B][D] */);
/* Policy::setParameters / getParameters (include/.../covariant_trajectory_policy.h:200-227). */
int stomp_engine_set_parameters(void* engine, const double* theta /* [B][D][N] */);
int stomp_engine_get_parameters(void* engine, double* theta /* [B][D][N] */);
/* Policy::updateParameters with row 0 of the updates (src/covariant_trajectory_policy.cpp:306-323). */
int stomp_engine_update_parameters(void* engine, const double* updates /* [B][D][N] */);
/* Policy::computeControlCosts(noise variant) for caller-supplied vectors
 * (src/covariant_trajectory_policy.cpp:228-255).  n = vectors per problem. */
int stomp_engine_compute_control_costs(void* engine, const double* parameters /* [B][n][D][N] */,
                                       const double* noise /* [B][n][D][N] */, int32_t n, double weight,
                                       double* control_costs /* [B][n][D][N] */);

/* ---- noise ---------------------------------------------------------------------- */
int stomp_engine_seed(void* engine, uint64_t seed);
/* Host-injection mode: eps[B][R_gen][D][N] used by the NEXT get_rollouts / iterate instead of the
 * engine RNG (already scaled by the noise stddev, i.e. Rollout::noise_).  n = rollouts per problem
 * supplied (>= the number generated that iteration). */
int stomp_engine_inject_noise(void* engine, const double* eps, int32_t n);
/* Same, but the host->device copy runs on the handle's copy stream and the call returns at once: the next get_rollouts /
 * iterate waits for it on the device.  Two device buffers alternate, so the noise of iteration i+1 can be uploaded
 * while iteration i computes.  eps must stay valid (and should be pinned) until the copy has been consumed or
 * stomp_engine_synchronize returns.  At most one injection can be pending. */
int stomp_engine_inject_noise_async(void* engine, const double* eps, int32_t n);
/* Draw the engine's own standard normals z[B][n][D][N] (iteration, global rollout id keyed Philox)
 * without touching engine state: statistical validation of the RNG against N(0, R^-1). */
int stomp_engine_sample_noise(void* engine, int32_t iteration, int32_t n, double* eps_out /* [B][n][D][N], unit stddev */);

/* ---- cost plugin: batched Task::execute ---------------------------------------- */
/* StompOptimizer::execute for n rollouts per problem (src/stomp_optimizer.cpp:1063-1165):
 * joint-limit projection, FK, collision spheres, SDF potential, FD velocity, cumulative obstacle cost.
 * collision_free may be NULL.  iteration_number == 1 reproduces the reference's first outer iteration
 * (iteration_ == 0), where the fixed start/goal padding points also count towards collision_free
 * (src/stomp_optimizer.cpp:624-630). */
int stomp_engine_execute(void* engine, const double* parameters /* [B][n][D][N] */, int32_t n,
                         int32_t iteration_number, double* costs /* [B][n][N] */,
                         int32_t* collision_free /* [B][n] */);
/* last_trajectory_constraints_satisfied_ of the rollouts of the last stomp_engine_execute call: [B][n]. */
int stomp_engine_execute_constraints_satisfied(void* engine, int32_t* satisfied, size_t count);
/* Same, plus the per-sphere records of performForwardKinematics (src/stomp_optimizer.cpp:618-709)
 * for rollout 0 of problem 0: debug[N][K]. */
int stomp_engine_execute_debug(void* engine, const double* parameters /* [D][N] */,
                               stomp_sphere_debug* debug /* [N+3][K]: trajectory points -1 .. N+1 */);

/* ---- PolicyImprovement, step by step (keeps the caller's own Task) ---------------- */
/* PolicyImprovement::getRollouts (src/policy_improvement.cpp:241-260).  Writes the newly generated
 * rollouts [B][R_gen][D][N] and R_gen. */
int stomp_engine_get_rollouts(void* engine, const double* noise_stddev /* [D] */, double* rollouts,
                              int32_t* num_generated);
/* PolicyImprovement::setRolloutCosts (src/policy_improvement.cpp:262-281).  costs[B][R_gen][N];
 * rollout_costs_total[B][R] may be NULL. */
int stomp_engine_set_rollout_costs(void* engine, const double* costs, double control_cost_weight,
                                   double* rollout_costs_total);
/* PolicyImprovement::improvePolicy (src/policy_improvement.cpp:385-401): updates[B][D][N] = row 0. */
int stomp_engine_improve_policy(void* engine, double* updates);
/* PolicyImprovement::addExtraRollouts with the current policy parameters as the single extra
 * rollout (src/policy_improvement.cpp:443-462; src/policy_improvement_loop.cpp:186-192). costs[B][N]. */
int stomp_engine_add_extra_rollouts(void* engine, const double* costs);

/* ---- the hot path --------------------------------------------------------------- */
/* PolicyImprovementLoop::runSingleIteration(iteration_number) with the built-in GPU Task
 * (src/policy_improvement_loop.cpp:143-202).  stats may be NULL (no host sync then). */
int stomp_engine_iterate(void* engine, int32_t iteration_number, stomp_iter_stats* stats);
/* iterations first..first+count-1 back to back on the device, no host round trip in between.  With graph mode 1, steady-state
 * iterations (rollout reuse primed, engine noise, no profiling) are replayed from a CUDA graph of 8 captured iterations of the
 * two-stream schedule: one graph launch instead of ~90 kernel launches, with the Philox generation counter and the
 * per-iteration noise scales advanced on the device; results are bit-identical to stomp_engine_iterate called count times. */
int stomp_engine_run(void* engine, int32_t first_iteration, int32_t count, stomp_iter_stats* last_stats);
/* 0 (default): stomp_engine_run launches every kernel from the host; 1: it replays CUDA graphs when it can.  Off by default
 * because the loop is bound by kernel-to-kernel dependency latency, not by host launch rate, and graph nodes resolve their
 * cross-branch dependencies more slowly than stream launches do (measured on B200: C1 0.140 vs 0.093 ms per iteration, C2 0.570
 * vs 0.505; profiles/README.md).  env STOMP_GRAPH sets the default. */
int stomp_engine_set_graph_mode(void* engine, int32_t mode);
int stomp_engine_synchronize(void* engine);
/* Results of the last iteration's noise-less rollout (what stomp_engine_iterate returns through `stats`), for callers
 * that launched the iteration with stats == NULL; waits for the iteration to finish. */
int stomp_engine_last_stats(void* engine, stomp_iter_stats* stats);

/* Asynchronous read-back of the results of the iteration launched last (stomp_engine_iterate with stats == NULL): the
 * updated trajectories theta [B][D][N], the noise-less rollout's cost [B] and collision flag [B] (any may be NULL) are
 * snapshotted on the device in stream order and copied to the caller's (pinned) host buffers on a dedicated stream, without
 * blocking the host or the iterations launched afterwards.  Two requests can be in flight; *ticket receives 0 or 1 and
 * stomp_engine_wait_results(ticket) blocks until that request's host buffers are complete. */
int stomp_engine_request_results_async(void* engine, double* theta, double* noiseless_cost, int32_t* noiseless_collision_free,
                                       int32_t* ticket);
int stomp_engine_wait_results(void* engine, int32_t ticket);

/* StompOptimizer::optimize, STOMP branch (src/stomp_optimizer.cpp:284-359,368-400), for every problem of the batch with
 * the bookkeeping on the device: per iteration the noise-less rollout's cost / collision flag update
 * collision_free_iteration_, the success iterations, the best (joint-limit-clipped) trajectory and
 * last_improvement_iteration_; a problem stops being tracked once it has been collision free for
 * max_iterations_after_collision_free consecutive iterations (the reference's early exit), and the call returns when
 * every problem has stopped or after max_iterations.  Mirrors msg/STOMPStatistics.msg.  All arrays are [B] and may be
 * NULL; costs is [max_iterations][B] (iteration-major), entries after a problem's exit are left untouched. */
typedef struct stomp_optimize_stats {
  int32_t* success;
  int32_t* success_iteration;            /* -1 if never collision free */
  int32_t* collision_success_iteration;  /* -1 if never collision free */
  int32_t* last_improvement_iteration;   /* -1 if the first iteration stayed the best */
  int32_t* iterations;                   /* iterations run before the problem's loop ended */
  double* best_cost;
  double* costs;
} stomp_optimize_stats;
int stomp_engine_optimize(void* engine, int32_t max_iterations, int32_t max_iterations_after_collision_free,
                          stomp_optimize_stats* stats);

/* ---- getters -------------------------------------------------------------------- */
/* Copies a field (see stomp_field) to a host buffer of `bytes` bytes (fp64 unless noted). */
int stomp_engine_get(void* engine, int32_t field, void* out, size_t bytes);
/* Number of CUDA kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t stomp_engine_launch_count(void* engine);
/* CUDA stream of the handle as an opaque pointer (for event timing on the launching stream). */
void* stomp_engine_stream(void* engine);
/* Device timing helpers on the handle's stream: record/elapsed of two internal events. */
int stomp_engine_timer_start(void* engine);
int stomp_engine_timer_stop(void* engine, float* elapsed_ms);

/* Per-kernel device timing (CUDA events on the handle's stream around every launch): enable, run, then sum
 * the durations of the launches whose kernel name contains kernel_substr ("" = all).  Used by bench.py for
 * the roofline of the dominant kernel; replaces the reference's commented-out ros::WallTime phase timers
 * (src/policy_improvement.cpp:255-257,389-397). */
int stomp_engine_set_profiling(void* engine, int32_t enabled);
int stomp_engine_get_profile(void* engine, const char* kernel_substr, double* total_ms, int64_t* num_launches);
/* enabled == 2 in stomp_engine_set_profiling keeps the two-stream schedule while recording; this writes what was recorded as
 * CSV (index, kernel, stream 0 main / 1 tail, begin_us, end_us after the first recorded launch).  Measurement aid only. */
int stomp_engine_dump_timeline(void* engine, const char* path);

/* ---- rollout sharding over GPUs (config C3) ----------------------------------------- */
/* Device buffers the host plumbing (torch.distributed / NCCL) all-reduces between the phases of
 * a sharded iteration: minmax = [2][D][N] (MAX of {c, -c}), sums = [2][D][N] (SUM of {e, e*eps}). */
int stomp_engine_shard_buffers(void* engine, void** minmax_dev, void** sums_dev, size_t* bytes_each);
int stomp_engine_iterate_sharded_phase(void* engine, int32_t iteration_number, int32_t phase /* 0,1,2 */);
/* The same iteration with both exchanges done by the GPUs themselves over NVLink peer memory, as the epilogue of the statistics
 * kernel (k_shard_stats: P2P stores of the 2*D*N partials into every peer's exchange buffer, system-scope flags, reduction in
 * rank order, then — SUM phase — the projection and theta += update in the same launch) instead of NCCL calls between
 * host-synchronised phases.  Setup, once: every rank exports its exchange buffer with stomp_engine_shard_ipc_handle (64-byte
 * cudaIpcMemHandle_t), the host side all-gathers the handles (any transport), every rank maps its peers with
 * stomp_engine_shard_open_peers(handles[world][64]) and the ranks barrier once before the first iteration (open_peers clears
 * this rank's flags and restarts the epochs, also when called again on a live engine).  stomp_engine_iterate_sharded_fused
 * then enqueues the whole iteration on the handle's stream without a host synchronisation; every rank must call it for the
 * same iteration.  An exchange that times out (a rank that never arrived) sets a sticky flag: the reduction is skipped, no
 * update is applied from then on, and stomp_engine_shard_status / stomp_engine_get_parameters / the iteration statistics
 * return the error. */
int stomp_engine_shard_ipc_handle(void* engine, void* handle_out, size_t handle_bytes);
int stomp_engine_shard_open_peers(void* engine, const void* handles, int32_t count);
int stomp_engine_iterate_sharded_fused(void* engine, int32_t iteration_number);
int stomp_engine_shard_status(void* engine);

#ifdef __cplusplus
}
#endif
#endif /* STOMP_B200_H_ */
