/*
 * stomp_oracle.cpp — CPU restatement of the STOMP per-iteration rollout loop of
 * kalakris/stomp_motion_planner_icra2011, single-threaded fp64, dependency-free C++17.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (the CUDA engine behind
 * include/stomp_b200.h) never links, imports or calls anything in oracle/.
 *
 * PARITY PINNED against outputs of the reference itself: the reference ships no unit tests, golden vectors or
 * known-answer data for this path and its rosbuild build cannot run here, but 13 of its 14
 * translation units compile UNMODIFIED from /root/reference against stand-in headers for the third-party packages it
 * does not vendor (oracle/ref_shim/, oracle/ref_driver.cpp -> oracle/_ref/libstomp_ref.so).  tests/golden/ref_*.npz are
 * that library's outputs; tests/test_reference_pinning.py holds this oracle to them (every intermediate of
 * runSingleIteration, every per-sphere quantity of StompOptimizer::execute, the statistics of optimize()): integer work
 * exact, floating point <= 1e-8.  What stays a restatement on BOTH sides are the un-vendored third-party semantics
 * listed below.
 *
 * The oracle keeps the reference's algorithmic shape on purpose (dense (N+12)^2
 * differentiation-matrix mat-vecs, dense L*z and M*eps, one serial execute per rollout), so
 * that its timing reflects the reference's cost structure.  Citations are relative to
 * /root/reference/stomp_motion_planner/.
 *
 * Third-party arithmetic that is not vendored in the reference and is restated from the
 * published behaviour of those packages:
 *   - orocos KDL Frame / Rotation::Rot2 / Segment::pose          (FK)
 *   - ROS distance_field::PropagationDistanceField lookup         (nearest cell, 1-cell margin)
 *   - Eigen 2 MatrixXd::inverse() (LU) and llt().matrixL()        (dense LA)
 */
#include "../include/stomp_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <utility>
#include <vector>

namespace {

typedef std::vector<double> Vec;

/* include/stomp_motion_planner/stomp_utils.h:49-56 */
const int DIFF_RULE_LENGTH = 7;
const int NUM_DIFF_RULES = 3;
const double DIFF_RULES[NUM_DIFF_RULES][DIFF_RULE_LENGTH] = {
    {0, 0, -2 / 6.0, -3 / 6.0, 6 / 6.0, -1 / 6.0, 0},
    {0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0},
    {0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0}};

struct Mat {
  int rows = 0, cols = 0;
  Vec a;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c), a(size_t(r) * c, 0.0) {}
  double& operator()(int i, int j) { return a[size_t(i) * cols + j]; }
  double operator()(int i, int j) const { return a[size_t(i) * cols + j]; }
};

Mat identity(int n, double s) {
  Mat m(n, n);
  for (int i = 0; i < n; ++i) m(i, i) = s;
  return m;
}

/* A^T A */
Mat gram(const Mat& A) {
  Mat G(A.cols, A.cols);
  for (int k = 0; k < A.rows; ++k)
    for (int i = 0; i < A.cols; ++i) {
      double aki = A(k, i);
      if (aki == 0.0) continue;
      for (int j = 0; j < A.cols; ++j) G(i, j) += aki * A(k, j);
    }
  return G;
}

Mat block(const Mat& A, int r0, int c0, int nr, int nc) {
  Mat B(nr, nc);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) B(i, j) = A(r0 + i, c0 + j);
  return B;
}

/* Eigen 2 MatrixXd::inverse(): LU decomposition.  Gauss-Jordan with partial pivoting. */
bool inverse(const Mat& A, Mat& out) {
  int n = A.rows;
  Mat W = A;
  out = identity(n, 1.0);
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(W(c, c));
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(W(r, c)) > best) best = std::fabs(W(r, c)), piv = r;
    if (best == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) std::swap(W(piv, j), W(c, j)), std::swap(out(piv, j), out(c, j));
    double inv = 1.0 / W(c, c);
    for (int j = 0; j < n; ++j) W(c, j) *= inv, out(c, j) *= inv;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      double f = W(r, c);
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) W(r, j) -= f * W(c, j), out(r, j) -= f * out(c, j);
    }
  }
  return true;
}

/* Eigen 2 llt().matrixL(): lower Cholesky factor. */
bool cholesky_lower(const Mat& A, Mat& L) {
  int n = A.rows;
  L = Mat(n, n);
  for (int j = 0; j < n; ++j) {
    double s = A(j, j);
    for (int k = 0; k < j; ++k) s -= L(j, k) * L(j, k);
    if (s <= 0.0) return false;
    L(j, j) = std::sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = A(i, j);
      for (int k = 0; k < j; ++k) t -= L(i, k) * L(j, k);
      L(i, j) = t / L(j, j);
    }
  }
  return true;
}

Vec matvec(const Mat& A, const Vec& x) {
  Vec y(A.rows, 0.0);
  for (int i = 0; i < A.rows; ++i) {
    double s = 0.0;
    const double* row = &A.a[size_t(i) * A.cols];
    for (int j = 0; j < A.cols; ++j) s += row[j] * x[j];
    y[i] = s;
  }
  return y;
}

/* ---- KDL restatement ------------------------------------------ */
struct Frame {
  double R[9];
  double p[3];
};

Frame frame_identity() {
  Frame f;
  std::memset(&f, 0, sizeof(f));
  f.R[0] = f.R[4] = f.R[8] = 1.0;
  return f;
}

/* KDL Frame*Frame = (R1 R2, R1 p2 + p1) */
Frame mul(const Frame& a, const Frame& b) {
  Frame c;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      c.R[i * 3 + j] = a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j] + a.R[i * 3 + 2] * b.R[6 + j];
    c.p[i] = a.R[i * 3] * b.p[0] + a.R[i * 3 + 1] * b.p[1] + a.R[i * 3 + 2] * b.p[2] + a.p[i];
  }
  return c;
}

Frame inv(const Frame& a) {
  Frame c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) c.p[i] = -(c.R[i * 3] * a.p[0] + c.R[i * 3 + 1] * a.p[1] + c.R[i * 3 + 2] * a.p[2]);
  return c;
}

void apply(const Frame& f, const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = f.R[i * 3] * v[0] + f.R[i * 3 + 1] * v[1] + f.R[i * 3 + 2] * v[2] + f.p[i];
}

/* KDL Rotation::Rot2(axis, angle), axis of unit length */
void rot2(const double v[3], double angle, double R[9]) {
  double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
  double m_vt_0 = vt * v[0], m_vt_1 = vt * v[1], m_vt_2 = vt * v[2];
  double m_st_0 = v[0] * st, m_st_1 = v[1] * st, m_st_2 = v[2] * st;
  double m_vt_0_1 = m_vt_0 * v[1], m_vt_0_2 = m_vt_0 * v[2], m_vt_1_2 = m_vt_1 * v[2];
  R[0] = ct + m_vt_0 * v[0];
  R[1] = -m_st_2 + m_vt_0_1;
  R[2] = m_st_1 + m_vt_0_2;
  R[3] = m_st_2 + m_vt_0_1;
  R[4] = ct + m_vt_1 * v[1];
  R[5] = -m_st_0 + m_vt_1_2;
  R[6] = -m_st_1 + m_vt_0_2;
  R[7] = m_st_0 + m_vt_1_2;
  R[8] = ct + m_vt_2 * v[2];
}

/* kdl_parser segment: pose(q) = Frame(Rot(axis,q)*rot, pos) | Frame(rot, pos+q*axis) | Frame(rot,pos) */
Frame segment_pose(const stomp_segment& s, double q) {
  Frame f;
  if (s.joint_type == STOMP_JOINT_REVOLUTE) {
    double Rq[9];
    rot2(s.axis, q, Rq);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        f.R[i * 3 + j] = Rq[i * 3] * s.rot[j] + Rq[i * 3 + 1] * s.rot[3 + j] + Rq[i * 3 + 2] * s.rot[6 + j];
    for (int i = 0; i < 3; ++i) f.p[i] = s.pos[i];
  } else {
    for (int i = 0; i < 9; ++i) f.R[i] = s.rot[i];
    for (int i = 0; i < 3; ++i) f.p[i] = s.pos[i] + (s.joint_type == STOMP_JOINT_PRISMATIC ? q * s.axis[i] : 0.0);
  }
  return f;
}

/* include/stomp_motion_planner/policy_improvement.h:50-63 */
struct Rollout {
  std::vector<Vec> parameters, noise, noise_projected, control_costs, total_costs, cumulative_costs, probabilities;
  Vec state_costs;
  double getCost() const { /* src/policy_improvement.cpp:149-156 */
    double cost = 0.0;
    for (double v : state_costs) cost += v;
    for (const Vec& c : control_costs)
      { double s = 0.0; for (double v : c) s += v; cost += s; }
    return cost;
  }
};

struct Oracle {
  stomp_engine_desc desc;
  int D = 0, N = 0, R = 0, Rreuse = 0, Nall = 0, fs = 0, fe = 0;
  double dt = 0;
  std::string err;

  /* robot / scene */
  std::vector<stomp_segment> segs;
  int ref_seg = 0;
  std::vector<stomp_sphere> spheres;
  std::vector<stomp_joint_limit> limits;
  int K = 0;
  std::vector<uint8_t> vox;
  int nx = 0, ny = 0, nz = 0, vox_dtype = 0;
  double origin[3] = {0, 0, 0}, res = 1.0;
  Vec sqrt_table;
  Vec noise_stddev, noise_decay;

  /* CovariantTrajectoryPolicy state (src/covariant_trajectory_policy.cpp) */
  Mat A[NUM_DIFF_RULES];
  Mat R_all, Rfree, Rinv;
  std::vector<Vec> params_all; /* [D][Nall] */

  /* PolicyImprovement state (src/policy_improvement.cpp) */
  Mat Lchol, Mproj;
  std::vector<Rollout> rollouts, reused_rollouts, extra_rollouts;
  std::vector<Vec> pi_parameters; /* parameters_ */
  std::vector<Vec> parameter_updates;
  int num_rollouts_gen = 0;
  bool rollouts_reused_next = false, extra_rollouts_added = false;
  double control_cost_weight = 0.0;
  std::mt19937 rng{1};
  std::normal_distribution<double> normal{0.0, 1.0};

  /* StompOptimizer cost-plugin state (src/stomp_optimizer.cpp) */
  Mat Qinv;
  std::vector<Vec> group_traj; /* [D][Nall], column-per-joint like the Eigen column-major MatrixXd */
  std::vector<double> cp_pos;  /* [Nall][K][3] */
  Vec cp_pot, cp_vel_mag;      /* [Nall][K] */
  std::vector<int> cp_coll, cp_vox; /* [Nall][K], [Nall][K][3] */
  std::vector<int> state_coll;
  int iteration_ = 0;
  bool last_collision_free = false;
  double last_cost = 0.0;
  Vec noiseless_costs;
  std::vector<int> collision_free_slots; /* [R+1] */
  /* inverse dynamics of the planning group's chain (src/stomp_optimizer.cpp:1006-1061; KDL::ChainIdSolver_RNE restated) */
  struct ChainLink { int seg; double m, h[3], I[9]; };
  std::vector<ChainLink> chain;
  double gravity[3] = {0, 0, -9.8};
  double torque_cost_weight = 0.0;
  Vec last_torques; /* [N][D] of the last execute (parity tap) */

  /* constraint evaluators (src/constraint_evaluator.cpp) */
  std::vector<stomp_orientation_constraint> constraints;
  double constraint_cost_weight = 0.0;
  std::vector<Frame> cp_frames;  /* [Nall][segments] */
  bool last_constraints_satisfied = true;
  std::vector<int> constraints_satisfied_slots; /* [R+1] */
};

/* ---- CovariantTrajectoryPolicy ------------------------------------------------------ */

/* createDifferentiationMatrices, src/covariant_trajectory_policy.cpp:204-226 */
void create_differentiation_matrices(Oracle& o) {
  double multiplier = 1.0;
  for (int d = 0; d < NUM_DIFF_RULES; ++d) {
    o.A[d] = Mat(o.Nall, o.Nall);
    multiplier /= o.dt;
    for (int i = 0; i < o.Nall; ++i)
      for (int j = -DIFF_RULE_LENGTH / 2; j <= DIFF_RULE_LENGTH / 2; ++j) {
        int index = i + j;
        if (index < 0 || index >= o.Nall) continue;
        o.A[d](i, index) = multiplier * DIFF_RULES[d][j + DIFF_RULE_LENGTH / 2];
      }
  }
}

/* initializeVariables + initializeCosts, src/covariant_trajectory_policy.cpp:150-191 */
bool policy_initialize(Oracle& o) {
  o.dt = o.desc.movement_duration / (o.N + 1);
  o.Nall = o.N + 2 * (DIFF_RULE_LENGTH - 1);
  o.fs = DIFF_RULE_LENGTH - 1;
  o.fe = o.fs + o.N - 1;
  o.params_all.assign(o.D, Vec(o.Nall, 0.0));
  create_differentiation_matrices(o);
  o.R_all = identity(o.Nall, o.desc.ridge_factor);
  for (int i = 0; i < NUM_DIFF_RULES; ++i) {
    if (o.desc.derivative_costs[i] == 0.0) continue; /* adds exact zeros otherwise */
    Mat G = gram(o.A[i]);
    for (size_t k = 0; k < G.a.size(); ++k) o.R_all.a[k] += o.desc.derivative_costs[i] * G.a[k];
  }
  o.Rfree = block(o.R_all, o.fs, o.fs, o.N, o.N);
  if (!inverse(o.Rfree, o.Rinv)) { o.err = "control cost matrix is singular"; return false; }
  return true;
}

/* setToMinControlCost / computeLinearControlCosts / computeMinControlCostParameters,
 * src/covariant_trajectory_policy.cpp:102-148 */
void policy_set_to_min_control_cost(Oracle& o, const double* start, const double* goal) {
  for (int d = 0; d < o.D; ++d) {
    for (int i = 0; i < DIFF_RULE_LENGTH - 1; ++i) {
      o.params_all[d][i] = start[d];
      o.params_all[d][o.Nall - 1 - i] = goal[d];
    }
    Vec lin(o.N, 0.0);
    for (int j = 0; j < o.N; ++j) {
      double s = 0.0;
      for (int i = 0; i < DIFF_RULE_LENGTH - 1; ++i) s += o.params_all[d][i] * o.R_all(i, o.fs + j);
      double s2 = 0.0;
      for (int i = 0; i < DIFF_RULE_LENGTH - 1; ++i) s2 += o.params_all[d][o.fe + 1 + i] * o.R_all(o.fe + 1 + i, o.fs + j);
      lin[j] = 2.0 * (s + s2);
    }
    Vec x = matvec(o.Rinv, lin);
    for (int j = 0; j < o.N; ++j) o.params_all[d][o.fs + j] = -0.5 * x[j];
  }
}

/* computeControlCosts (noise variant), src/covariant_trajectory_policy.cpp:228-255 */
void policy_compute_control_costs(const Oracle& o, const std::vector<Vec>& parameters, const std::vector<Vec>& noise,
                                  double weight, std::vector<Vec>& control_costs) {
  for (int d = 0; d < o.D; ++d) {
    Vec params_all = o.params_all[d];
    Vec costs_all(o.Nall, 0.0);
    for (int j = 0; j < o.N; ++j) params_all[o.fs + j] = parameters[d][j] + noise[d][j];
    for (int i = 0; i < NUM_DIFF_RULES; ++i) {
      Vec acc_all = matvec(o.A[i], params_all); /* dense (N+12)^2 on purpose */
      for (int k = 0; k < o.Nall; ++k) costs_all[k] += weight * o.desc.derivative_costs[i] * (acc_all[k] * acc_all[k]);
    }
    control_costs[d].assign(costs_all.begin() + o.fs, costs_all.begin() + o.fs + o.N);
    for (int i = 0; i < o.fs; ++i) {
      control_costs[d][0] += costs_all[i];
      control_costs[d][o.N - 1] += costs_all[o.Nall - (i + 1)];
    }
  }
}

void policy_get_parameters(const Oracle& o, std::vector<Vec>& p) {
  p.resize(o.D);
  for (int d = 0; d < o.D; ++d) p[d].assign(o.params_all[d].begin() + o.fs, o.params_all[d].begin() + o.fs + o.N);
}

/* ---- StompCost (src/stomp_cost.cpp:47-105) + scaling (src/stomp_optimizer.cpp:105-125) -- */
bool build_quad_cost_inv(Oracle& o) {
  Mat full(o.Nall, o.Nall);
  double multiplier = 1.0;
  double smooth[3] = {o.desc.derivative_costs[0], o.desc.derivative_costs[1], o.desc.derivative_costs[2]};
  for (int i = 0; i < NUM_DIFF_RULES; ++i) {
    multiplier *= o.desc.discretization;
    Mat diff(o.Nall, o.Nall);
    for (int r = 0; r < o.Nall; ++r)
      for (int j = -DIFF_RULE_LENGTH / 2; j <= DIFF_RULE_LENGTH / 2; ++j) {
        int index = r + j;
        if (index < 0 || index >= o.Nall) continue;
        diff(r, index) = DIFF_RULES[i][j + DIFF_RULE_LENGTH / 2];
      }
    Mat G = gram(diff);
    for (size_t k = 0; k < G.a.size(); ++k) full.a[k] += (smooth[i] * multiplier) * G.a[k];
  }
  for (int i = 0; i < o.Nall; ++i) full(i, i) += o.desc.ridge_factor;
  Mat q = block(full, o.fs, o.fs, o.N, o.N);
  if (!inverse(q, o.Qinv)) { o.err = "quad cost matrix is singular"; return false; }
  /* all joint_costs are 1.0 -> every joint shares one matrix; scale by the global max coefficient */
  double mx = o.Qinv.a[0];
  for (double v : o.Qinv.a) mx = std::max(mx, v);
  double inv_scale = 1.0 / mx;
  for (double& v : o.Qinv.a) v *= inv_scale;
  return true;
}

/* ---- distance field lookup ------------------------------------- */
static double cell_distance(const Oracle& o, int cx, int cy, int cz) { /* a cell's distance incl. the one-cell margin rule */
  if (cx < 1 || cy < 1 || cz < 1 || cx >= o.nx - 1 || cy >= o.ny - 1 || cz >= o.nz - 1) return 0.0;
  size_t idx = (size_t(cx) * o.ny + cy) * o.nz + cz;
  switch (o.vox_dtype) {
    case STOMP_VOXEL_U8_SQ: return o.sqrt_table[o.vox[idx]];
    case STOMP_VOXEL_U16_SQ: return o.sqrt_table[reinterpret_cast<const uint16_t*>(o.vox.data())[idx]];
    default: return double(reinterpret_cast<const float*>(o.vox.data())[idx]);
  }
}

/* STOMP_SDF_TRILINEAR (engine extension, not a reference mode): trilinear interpolation of the eight cell distances around the
 * point, every corner following the nearest-cell rule above (0 on / outside the outermost layer), so the field is the
 * continuous extension of the reference's piecewise-constant one.  cell = the lower corner. */
double sdf_distance_trilinear(const Oracle& o, const double pos[3], int cell[3]) {
  double f[3];
  for (int i = 0; i < 3; ++i) {
    const double t = (pos[i] - o.origin[i]) / o.res;
    const double fl = std::floor(t);
    cell[i] = std::fabs(fl) < 2.0e9 ? int(fl) : -1;
    f[i] = t - fl;
  }
  double acc = 0.0;
  for (int dx = 0; dx < 2; ++dx)
    for (int dy = 0; dy < 2; ++dy)
      for (int dz = 0; dz < 2; ++dz) {
        const double w = (dx ? f[0] : 1.0 - f[0]) * (dy ? f[1] : 1.0 - f[1]) * (dz ? f[2] : 1.0 - f[2]);
        acc += w * cell_distance(o, cell[0] + dx, cell[1] + dy, cell[2] + dz);
      }
  return acc;
}

double sdf_distance(const Oracle& o, const double pos[3], int cell[3]) {
  if (o.desc.sdf_mode == STOMP_SDF_TRILINEAR) return sdf_distance_trilinear(o, pos, cell);
  const int n[3] = {o.nx, o.ny, o.nz};
  bool outside = false;
  for (int i = 0; i < 3; ++i) {
    cell[i] = int(std::round((pos[i] - o.origin[i]) / o.res));
    if (cell[i] < 1 || cell[i] >= n[i] - 1) outside = true;
  }
  if (outside) return 0.0;
  size_t idx = (size_t(cell[0]) * o.ny + cell[1]) * o.nz + cell[2];
  switch (o.vox_dtype) {
    case STOMP_VOXEL_U8_SQ: return o.sqrt_table[o.vox[idx]];
    case STOMP_VOXEL_U16_SQ: return o.sqrt_table[reinterpret_cast<const uint16_t*>(o.vox.data())[idx]];
    default: return double(reinterpret_cast<const float*>(o.vox.data())[idx]);
  }
}

/* getCollisionPointPotentialGradient, include/stomp_motion_planner/stomp_collision_space.h:193-228
 * (the gradient is not used by the STOMP cost and is not restated) */
bool collision_potential(const Oracle& o, const stomp_sphere& s, const double pos[3], double& potential, int cell[3]) {
  double field_distance = sdf_distance(o, pos, cell);
  double d = field_distance - s.radius;
  if (d >= s.clearance) {
    potential = 0.0;
  } else if (d >= 0.0) {
    double diff = d - s.clearance;
    double gradient_magnitude = diff * (1.0 / s.clearance);
    potential = 0.5 * gradient_magnitude * diff;
  } else {
    potential = -d + 0.5 * s.clearance;
  }
  return field_distance <= s.radius;
}

/* ---- OrientationConstraintEvaluator: src/constraint_evaluator.cpp:50-114 (bullet's btMatrix3x3 restated) ---- */
void bt_set_rotation(const double q[4], double m[9]) { /* btMatrix3x3::setRotation(btQuaternion(x,y,z,w)) */
  double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double s = 2.0 / d;
  double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
  double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
  double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  m[0] = 1.0 - (yy + zz); m[1] = xy - wz; m[2] = xz + wy;
  m[3] = xy + wz; m[4] = 1.0 - (xx + zz); m[5] = yz - wx;
  m[6] = xz - wy; m[7] = yz + wx; m[8] = 1.0 - (xx + yy);
}
void bt_get_rpy(const double m[9], double& roll, double& pitch, double& yaw) { /* btMatrix3x3::getEulerYPR, solution 1 */
  if (std::fabs(m[6]) >= 1.0) {
    yaw = 0.0;
    double delta = std::atan2(m[7], m[8]);
    if (m[6] < 0.0) { pitch = M_PI / 2.0; roll = delta; }
    else { pitch = -M_PI / 2.0; roll = delta; }
  } else {
    pitch = -std::asin(m[6]);
    double cp = std::cos(pitch);
    roll = std::atan2(m[7] / cp, m[8] / cp);
    yaw = std::atan2(m[3] / cp, m[0] / cp);
  }
}
bool constraint_cost(const stomp_orientation_constraint& c, const Frame& f, double& cost) {
  double nom[9], inv[9], res[9];
  bt_set_rotation(c.orientation, nom);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inv[i * 3 + j] = nom[j * 3 + i]; /* inverse of a rotation */
  const double* a = c.body_fixed ? inv : f.R;
  const double* b = c.body_fixed ? f.R : inv;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) res[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
  double roll, pitch, yaw;
  bt_get_rpy(res, roll, pitch, yaw);
  roll = std::fabs(roll); pitch = std::fabs(pitch); yaw = std::fabs(yaw);
  double rw = c.absolute_roll_tolerance >= M_PI ? 0.0 : 1.0, pw = c.absolute_pitch_tolerance >= M_PI ? 0.0 : 1.0,
         yw = c.absolute_yaw_tolerance >= M_PI ? 0.0 : 1.0;
  cost = c.weight * (rw * roll + pw * pitch + yw * yaw);
  return !(roll > c.absolute_roll_tolerance || pitch > c.absolute_pitch_tolerance || yaw > c.absolute_yaw_tolerance);
}

/* ---- forward kinematics: src/treefksolverjointposaxis_partial.cpp:76-178 -------------- */
void forward_kinematics(const Oracle& o, const double* q_group, std::vector<Frame>& frames) {
  int S = int(o.segs.size());
  frames.resize(S);
  for (int s = 0; s < S; ++s) { /* segments are in DFS pre-order: parent index < own index */
    const stomp_segment& sg = o.segs[s];
    double q = sg.group_index >= 0 ? q_group[sg.group_index] : sg.fixed_value;
    Frame parent = sg.parent >= 0 ? frames[sg.parent] : frame_identity();
    frames[s] = mul(parent, segment_pose(sg, q));
  }
  Frame inv_ref = inv(frames[o.ref_seg]);
  for (int s = 0; s < S; ++s) frames[s] = mul(inv_ref, frames[s]);
}

/* handleJointLimits, src/stomp_optimizer.cpp:562-616 */
void handle_joint_limits(Oracle& o) {
  for (int joint = 0; joint < o.D; ++joint) {
    if (!o.limits[joint].has_limits) continue;
    double joint_max = o.limits[joint].max, joint_min = o.limits[joint].min;
    int count = 0;
    bool violation = false;
    do {
      double max_abs_violation = 1e-6, max_violation = 0.0;
      int max_violation_index = 0;
      violation = false;
      for (int i = o.fs; i <= o.fe; ++i) {
        double amount = 0.0, absolute_amount = 0.0;
        double v = o.group_traj[joint][i];
        if (v > joint_max) { amount = joint_max - v; absolute_amount = std::fabs(amount); }
        else if (v < joint_min) { amount = joint_min - v; absolute_amount = std::fabs(amount); }
        if (absolute_amount > max_abs_violation) {
          max_abs_violation = absolute_amount;
          max_violation = amount;
          max_violation_index = i;
          violation = true;
        }
      }
      if (violation) {
        int fv = max_violation_index - o.fs;
        double multiplier = max_violation / o.Qinv(fv, fv);
        for (int i = 0; i < o.N; ++i) o.group_traj[joint][o.fs + i] += multiplier * o.Qinv(i, fv);
      }
      if (++count > 10) break;
    } while (violation);
  }
}

/* performForwardKinematics, src/stomp_optimizer.cpp:618-709 */
bool perform_forward_kinematics(Oracle& o) {
  double invTime = 1.0 / o.desc.discretization;
  int start = o.fs, end = o.fe;
  if (o.iteration_ == 0) { start = 0; end = o.Nall - 1; }
  bool is_collision_free = true;
  std::vector<Frame> frames;
  Vec q(o.D);
  for (int i = start; i <= end; ++i) {
    for (int d = 0; d < o.D; ++d) q[d] = o.group_traj[d][i];
    forward_kinematics(o, q.data(), frames);
    if (!o.constraints.empty()) std::copy(frames.begin(), frames.end(), o.cp_frames.begin() + size_t(i) * o.segs.size());
    o.state_coll[i] = 0;
    for (int j = 0; j < o.K; ++j) {
      double* pos = &o.cp_pos[(size_t(i) * o.K + j) * 3];
      apply(frames[o.spheres[j].segment], o.spheres[j].pos, pos); /* stomp_collision_point.h:138-141 */
      bool colliding = collision_potential(o, o.spheres[j], pos, o.cp_pot[size_t(i) * o.K + j], &o.cp_vox[(size_t(i) * o.K + j) * 3]);
      o.cp_coll[size_t(i) * o.K + j] = colliding;
      if (colliding) o.state_coll[i] = 1;
    }
    if (o.state_coll[i]) is_collision_free = false;
  }
  for (int i = o.fs; i <= o.fe; ++i)
    for (int j = 0; j < o.K; ++j) {
      double vel[3] = {0, 0, 0};
      for (int k = -DIFF_RULE_LENGTH / 2; k <= DIFF_RULE_LENGTH / 2; ++k) {
        double c = invTime * DIFF_RULES[0][k + DIFF_RULE_LENGTH / 2];
        const double* p = &o.cp_pos[(size_t(i + k) * o.K + j) * 3];
        vel[0] += c * p[0]; vel[1] += c * p[1]; vel[2] += c * p[2];
      }
      o.cp_vel_mag[size_t(i) * o.K + j] = std::sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
    }
  return is_collision_free;
}

/* ---- inverse dynamics: KDL::ChainIdSolver_RNE::CartToJnt restated (orocos KDL 1.0, Featherstone's recursive Newton-Euler in
 * segment-tip coordinates; spatial vectors kept as (linear, angular) pairs like KDL::Twist / KDL::Wrench) ------------------ */
inline void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void rot_t(const double R[9], const double v[3], double o[3]) { /* R^T v */
  for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
inline void rot_n(const double R[9], const double v[3], double o[3]) { /* R v */
  for (int i = 0; i < 3; ++i) o[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}
struct Spatial { double lin[3], ang[3]; };
/* Frame::Inverse(Twist): the twist seen from the child frame X */
Spatial twist_to_child(const Frame& X, const Spatial& t) {
  double pxw[3], d[3];
  Spatial o;
  cross3(X.p, t.ang, pxw);
  for (int i = 0; i < 3; ++i) d[i] = t.lin[i] - pxw[i];
  rot_t(X.R, d, o.lin);
  rot_t(X.R, t.ang, o.ang);
  return o;
}
/* RigidBodyInertia * Twist */
Spatial inertia_times(const Oracle::ChainLink& L, const Spatial& t) {
  Spatial w;
  double hxw[3], hxv[3], Iw[3];
  cross3(L.h, t.ang, hxw);
  cross3(L.h, t.lin, hxv);
  rot_n(L.I, t.ang, Iw);
  for (int i = 0; i < 3; ++i) { w.lin[i] = L.m * t.lin[i] - hxw[i]; w.ang[i] = Iw[i] + hxv[i]; }
  return w;
}
/* torques[D] of one trajectory point (group joint values q, velocities qd, accelerations qdd) */
void chain_inverse_dynamics(const Oracle& o, const double* q, const double* qd, const double* qdd, double* torques) {
  const int ns = int(o.chain.size());
  std::vector<Frame> X(ns);
  std::vector<Spatial> S(ns), f(ns);
  Spatial v = {}, a = {};
  for (int i = 0; i < ns; ++i) {
    const stomp_segment& sg = o.segs[o.chain[i].seg];
    const int j = sg.group_index;
    const double q_ = j >= 0 ? q[j] : 0.0, qd_ = j >= 0 ? qd[j] : 0.0, qdd_ = j >= 0 ? qdd[j] : 0.0;
    X[i] = segment_pose(sg, q_);
    /* unit twist of the joint in the parent frame: the segment tip sits on the joint origin (kdl_parser), so no RefPoint shift */
    Spatial unit = {};
    if (j >= 0 && sg.joint_type == STOMP_JOINT_REVOLUTE) for (int k = 0; k < 3; ++k) unit.ang[k] = sg.axis[k];
    if (j >= 0 && sg.joint_type == STOMP_JOINT_PRISMATIC) for (int k = 0; k < 3; ++k) unit.lin[k] = sg.axis[k];
    rot_t(X[i].R, unit.lin, S[i].lin);
    rot_t(X[i].R, unit.ang, S[i].ang);
    Spatial vj;
    for (int k = 0; k < 3; ++k) { vj.lin[k] = S[i].lin[k] * qd_; vj.ang[k] = S[i].ang[k] * qd_; }
    Spatial vp, ap;
    if (i == 0) {
      vp = Spatial{};
      Spatial ag = {};
      for (int k = 0; k < 3; ++k) ag.lin[k] = -o.gravity[k];
      ap = twist_to_child(X[i], ag);
    } else {
      vp = twist_to_child(X[i], v);
      ap = twist_to_child(X[i], a);
    }
    for (int k = 0; k < 3; ++k) { v.lin[k] = vp.lin[k] + vj.lin[k]; v.ang[k] = vp.ang[k] + vj.ang[k]; }
    /* v x vj (Twist * Twist) */
    double c1[3], c2[3], c3[3];
    cross3(v.ang, vj.lin, c1);
    cross3(v.lin, vj.ang, c2);
    cross3(v.ang, vj.ang, c3);
    for (int k = 0; k < 3; ++k) {
      a.lin[k] = ap.lin[k] + S[i].lin[k] * qdd_ + c1[k] + c2[k];
      a.ang[k] = ap.ang[k] + S[i].ang[k] * qdd_ + c3[k];
    }
    /* f = I a + v x* (I v) */
    Spatial Ia = inertia_times(o.chain[i], a), Iv = inertia_times(o.chain[i], v);
    double d1[3], d2[3], d3[3];
    cross3(v.ang, Iv.lin, d1);
    cross3(v.ang, Iv.ang, d2);
    cross3(v.lin, Iv.lin, d3);
    for (int k = 0; k < 3; ++k) { f[i].lin[k] = Ia.lin[k] + d1[k]; f[i].ang[k] = Ia.ang[k] + d2[k] + d3[k]; }
  }
  for (int i = ns - 1; i >= 0; --i) {
    const int j = o.segs[o.chain[i].seg].group_index;
    if (j >= 0) {
      double t = 0.0;
      for (int k = 0; k < 3; ++k) t += S[i].lin[k] * f[i].lin[k] + S[i].ang[k] * f[i].ang[k];
      torques[j] = t;
    }
    if (i != 0) { /* Frame * Wrench into the parent's coordinates */
      double F[3], T[3], pxF[3];
      rot_n(X[i].R, f[i].lin, F);
      rot_n(X[i].R, f[i].ang, T);
      cross3(X[i].p, F, pxF);
      for (int k = 0; k < 3; ++k) { f[i - 1].lin[k] += F[k]; f[i - 1].ang[k] += T[k] + pxF[k]; }
    }
  }
}

/* StompOptimizer::getTorques, src/stomp_optimizer.cpp:1034-1060: joint velocities / accelerations by the 7-tap rules over the
 * group trajectory (include/stomp_motion_planner/stomp_trajectory.h:286-310), then the chain's inverse dynamics */
void get_torques(const Oracle& o, int index, double* torques) {
  Vec q(o.D), qd(o.D, 0.0), qdd(o.D, 0.0);
  const double invTime = 1.0 / o.desc.discretization, invTime2 = 1.0 / (o.desc.discretization * o.desc.discretization);
  for (int d = 0; d < o.D; ++d) {
    q[d] = o.group_traj[d][index];
    for (int k = -DIFF_RULE_LENGTH / 2; k <= DIFF_RULE_LENGTH / 2; ++k) {
      qd[d] += (invTime * DIFF_RULES[0][k + DIFF_RULE_LENGTH / 2]) * o.group_traj[d][index + k];
      qdd[d] += (invTime2 * DIFF_RULES[1][k + DIFF_RULE_LENGTH / 2]) * o.group_traj[d][index + k];
    }
  }
  chain_inverse_dynamics(o, q.data(), qd.data(), qdd.data(), torques);
}

/* StompOptimizer::execute, src/stomp_optimizer.cpp:1063-1165 */
void task_execute(Oracle& o, const std::vector<Vec>& parameters, Vec& costs, int iteration_number) {
  o.iteration_ = iteration_number - 1; /* optimize() calls runSingleIteration(iteration_+1) */
  for (int d = 0; d < o.D; ++d)
    for (int i = 0; i < o.N; ++i) o.group_traj[d][o.fs + i] = parameters[d][i];
  handle_joint_limits(o);
  if (!o.constraints.empty()) o.cp_frames.resize(size_t(o.Nall) * o.segs.size());
  o.last_collision_free = perform_forward_kinematics(o);
  o.last_constraints_satisfied = true;
  costs.assign(o.N, 0.0);
  o.last_torques.assign(size_t(o.N) * o.D, 0.0);
  for (int i = o.fs; i <= o.fe; ++i) {
    double state_collision_cost = 0.0, cumulative = 0.0;
    for (int j = 0; j < o.K; ++j) {
      cumulative += o.cp_pot[size_t(i) * o.K + j] * o.cp_vel_mag[size_t(i) * o.K + j];
      state_collision_cost += cumulative;
    }
    double state_constraint_cost = 0.0;
    for (const stomp_orientation_constraint& c : o.constraints) {
      double cost;
      if (!constraint_cost(c, o.cp_frames[size_t(i) * o.segs.size() + c.segment], cost)) o.last_constraints_satisfied = false;
      state_constraint_cost += cost;
    }
    double state_torque_cost = 0.0;
    if (o.torque_cost_weight > 1e-9) {
      Vec torques(o.D, 0.0);
      get_torques(o, i, torques.data());
      for (int j = 0; j < o.D; ++j) {
        state_torque_cost += std::fabs(torques[j]);
        o.last_torques[size_t(i - o.fs) * o.D + j] = torques[j];
      }
    }
    costs[i - o.fs] = o.desc.obstacle_cost_weight * state_collision_cost + o.constraint_cost_weight * state_constraint_cost +
                      o.torque_cost_weight * state_torque_cost;
  }
  double s = 0.0;
  for (double c : costs) s += c;
  o.last_cost = s;
}

/* ---- PolicyImprovement --------------------------------------------------------------- */

/* initialize + setNumRollouts + preComputeProjectionMatrices, src/policy_improvement.cpp:64-147,421-441 */
bool pi_initialize(Oracle& o) {
  policy_get_parameters(o, o.pi_parameters);
  if (!cholesky_lower(o.Rinv, o.Lchol)) { o.err = "R^-1 is not positive definite"; return false; }
  if (o.Rreuse >= o.R) { o.err = "Number of reused rollouts must be strictly less than number of rollouts."; return false; }
  Rollout r;
  r.parameters.assign(o.D, Vec(o.N, 0.0));
  r.noise = r.noise_projected = r.control_costs = r.total_costs = r.cumulative_costs = r.probabilities = r.parameters;
  r.state_costs.assign(o.N, 0.0);
  o.rollouts.assign(o.R, r);
  o.reused_rollouts.assign(o.Rreuse, r);
  o.extra_rollouts.assign(1, r);
  o.rollouts_reused_next = false;
  o.extra_rollouts_added = false;
  o.num_rollouts_gen = 0;
  o.parameter_updates.assign(o.D, Vec(o.N, 0.0));
  o.Mproj = o.Rinv;
  for (int p = 0; p < o.N; ++p) {
    double column_max = o.Rinv(0, p);
    for (int p2 = 1; p2 < o.N; ++p2)
      if (o.Rinv(p2, p) > column_max) column_max = o.Rinv(p2, p);
    double s = 1.0 / (o.N * column_max);
    for (int p2 = 0; p2 < o.N; ++p2) o.Mproj(p2, p) *= s;
  }
  return true;
}

/* computeProjectedNoise(Rollout&), src/policy_improvement.cpp:473-482 */
void pi_projected_noise(const Oracle& o, Rollout& r) {
  for (int d = 0; d < o.D; ++d) r.noise_projected[d] = matvec(o.Mproj, r.noise[d]);
}

/* computeRolloutControlCosts(Rollout&), src/policy_improvement.cpp:484-489 */
void pi_control_costs(const Oracle& o, Rollout& r) {
  policy_compute_control_costs(o, r.parameters, r.noise_projected, 0.5 * o.control_cost_weight, r.control_costs);
}

/* generateRollouts + getRollouts, src/policy_improvement.cpp:158-260.
 * eps_injected: [num_gen][D][N] already scaled noise (Rollout::noise_), or NULL -> MultivariateGaussian
 * (include/stomp_motion_planner/multivariate_gaussian.h:88-94) with a std::mt19937. */
void pi_get_rollouts(Oracle& o, const double* noise_stddev, const double* eps_injected) {
  policy_get_parameters(o, o.pi_parameters);
  o.num_rollouts_gen = o.R - o.Rreuse;
  if (!o.rollouts_reused_next) {
    o.num_rollouts_gen = o.R;
    if (o.Rreuse > 0) o.rollouts_reused_next = true;
  } else {
    std::vector<std::pair<double, int> > sorter;
    for (int r = 0; r < o.R; ++r) sorter.push_back(std::make_pair(o.rollouts[r].getCost(), r));
    if (o.extra_rollouts_added) {
      sorter.push_back(std::make_pair(o.extra_rollouts[0].getCost(), -1));
      o.extra_rollouts_added = false;
    }
    std::sort(sorter.begin(), sorter.end());
    for (int r = 0; r < o.Rreuse; ++r) {
      int reuse_index = sorter[r].second;
      o.reused_rollouts[r] = reuse_index >= 0 ? o.rollouts[reuse_index] : o.extra_rollouts[-reuse_index - 1];
    }
    for (int r = 0; r < o.Rreuse; ++r) {
      Rollout& dst = o.rollouts[o.num_rollouts_gen + r];
      dst = o.reused_rollouts[r];
      for (int d = 0; d < o.D; ++d)
        for (int i = 0; i < o.N; ++i) dst.noise[d][i] = dst.parameters[d][i] - o.pi_parameters[d][i];
    }
  }
  Vec z(o.N);
  for (int d = 0; d < o.D; ++d)
    for (int r = 0; r < o.num_rollouts_gen; ++r) {
      Rollout& ro = o.rollouts[r];
      if (eps_injected) {
        const double* e = eps_injected + (size_t(r) * o.D + d) * o.N;
        for (int i = 0; i < o.N; ++i) ro.noise[d][i] = e[i];
      } else {
        for (int i = 0; i < o.N; ++i) z[i] = o.normal(o.rng);
        Vec s = matvec(o.Lchol, z); /* dense L, like covariance_cholesky_*output */
        for (int i = 0; i < o.N; ++i) ro.noise[d][i] = noise_stddev[d] * s[i];
      }
      for (int i = 0; i < o.N; ++i) ro.parameters[d][i] = o.pi_parameters[d][i] + ro.noise[d][i];
    }
  for (int r = 0; r < o.R; ++r) pi_projected_noise(o, o.rollouts[r]);
}

/* setRolloutCosts, src/policy_improvement.cpp:262-281.  costs: [num_gen][N] */
void pi_set_rollout_costs(Oracle& o, const double* costs, double control_cost_weight, double* totals) {
  o.control_cost_weight = control_cost_weight;
  for (int r = 0; r < o.R; ++r) pi_control_costs(o, o.rollouts[r]);
  for (int r = 0; r < o.num_rollouts_gen; ++r)
    o.rollouts[r].state_costs.assign(costs + size_t(r) * o.N, costs + size_t(r + 1) * o.N);
  if (totals)
    for (int r = 0; r < o.R; ++r) totals[r] = o.rollouts[r].getCost();
}

/* improvePolicy, src/policy_improvement.cpp:301-401 */
void pi_improve_policy(Oracle& o) {
  for (int r = 0; r < o.R; ++r)
    for (int d = 0; d < o.D; ++d) {
      Rollout& ro = o.rollouts[r];
      for (int t = 0; t < o.N; ++t) ro.total_costs[d][t] = ro.state_costs[t] + ro.control_costs[d][t];
      ro.cumulative_costs[d] = ro.total_costs[d];
      if (o.desc.use_cumulative_costs)
        for (int t = o.N - 2; t >= 0; --t) ro.cumulative_costs[d][t] += ro.cumulative_costs[d][t + 1];
    }
  for (int d = 0; d < o.D; ++d)
    for (int t = 0; t < o.N; ++t) {
      double min_cost = o.rollouts[0].cumulative_costs[d][t], max_cost = min_cost;
      for (int r = 1; r < o.R; ++r) {
        double c = o.rollouts[r].cumulative_costs[d][t];
        if (c < min_cost) min_cost = c;
        if (c > max_cost) max_cost = c;
      }
      double denom = max_cost - min_cost;
      if (denom < 1e-8) denom = 1e-8;
      double p_sum = 0.0;
      for (int r = 0; r < o.R; ++r) {
        o.rollouts[r].probabilities[d][t] = std::exp(-10.0 * (o.rollouts[r].cumulative_costs[d][t] - min_cost) / denom);
        p_sum += o.rollouts[r].probabilities[d][t];
      }
      for (int r = 0; r < o.R; ++r) o.rollouts[r].probabilities[d][t] /= p_sum;
    }
  for (int d = 0; d < o.D; ++d) {
    Vec u(o.N, 0.0);
    for (int r = 0; r < o.R; ++r)
      for (int t = 0; t < o.N; ++t) u[t] += o.rollouts[r].noise[d][t] * o.rollouts[r].probabilities[d][t];
    o.parameter_updates[d] = matvec(o.Mproj, u);
  }
}

/* CovariantTrajectoryPolicy::updateParameters, src/covariant_trajectory_policy.cpp:306-323 */
void policy_update_parameters(Oracle& o, const std::vector<Vec>& updates) {
  for (int d = 0; d < o.D; ++d)
    for (int t = 0; t < o.N; ++t) o.params_all[d][o.fs + t] += 1.0 * updates[d][t];
}

/* addExtraRollouts, src/policy_improvement.cpp:443-471 */
void pi_add_extra_rollout(Oracle& o, const std::vector<Vec>& params, const Vec& costs) {
  policy_get_parameters(o, o.pi_parameters);
  Rollout& e = o.extra_rollouts[0];
  e.parameters = params;
  e.state_costs = costs;
  for (int d = 0; d < o.D; ++d)
    for (int i = 0; i < o.N; ++i) e.noise[d][i] = e.parameters[d][i] - o.pi_parameters[d][i];
  pi_projected_noise(o, e);
  pi_control_costs(o, e);
  o.extra_rollouts_added = true;
}

/* PolicyImprovementLoop::runSingleIteration, src/policy_improvement_loop.cpp:143-202 */
void run_single_iteration(Oracle& o, int iteration_number, const double* eps_injected) {
  Vec noise(o.D);
  for (int i = 0; i < o.D; ++i) noise[i] = o.noise_stddev[i] * std::pow(o.noise_decay[i], iteration_number - 1);
  pi_get_rollouts(o, noise.data(), eps_injected);
  Vec rollout_costs(size_t(o.num_rollouts_gen) * o.N), tmp;
  for (int r = 0; r < o.num_rollouts_gen; ++r) {
    task_execute(o, o.rollouts[r].parameters, tmp, iteration_number);
    std::copy(tmp.begin(), tmp.end(), rollout_costs.begin() + size_t(r) * o.N);
    o.collision_free_slots[r] = o.last_collision_free;
    o.constraints_satisfied_slots[r] = o.last_constraints_satisfied;
  }
  pi_set_rollout_costs(o, rollout_costs.data(), o.desc.smoothness_cost_weight, nullptr);
  pi_improve_policy(o);
  policy_update_parameters(o, o.parameter_updates);
  std::vector<Vec> theta;
  policy_get_parameters(o, theta);
  task_execute(o, theta, o.noiseless_costs, iteration_number);
  o.collision_free_slots[o.R] = o.last_collision_free;
  o.constraints_satisfied_slots[o.R] = o.last_constraints_satisfied;
  pi_add_extra_rollout(o, theta, o.noiseless_costs);
}

void flatten(const std::vector<Vec>& v, double* out) {
  for (size_t d = 0; d < v.size(); ++d) std::copy(v[d].begin(), v[d].end(), out + d * v[d].size());
}

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

} // namespace

extern "C" {

const char* stomp_oracle_last_error(void) { return g_err.c_str(); }

int stomp_oracle_create(const stomp_engine_desc* desc, void** out) {
  if (!desc || !out) return fail("null argument");
  Oracle* o = new Oracle();
  o->desc = *desc;
  o->D = desc->num_dimensions;
  o->N = desc->num_time_steps;
  o->R = desc->num_rollouts;
  o->Rreuse = desc->num_reused_rollouts;
  if (o->D < 1 || o->N < 1 || o->R < 1) { delete o; return fail("bad dimensions"); }
  if (!policy_initialize(*o) || !build_quad_cost_inv(*o) || !pi_initialize(*o)) {
    g_err = o->err;
    delete o;
    return 1;
  }
  o->group_traj.assign(o->D, Vec(o->Nall, 0.0));
  o->noise_stddev.assign(o->D, 1.0);
  o->noise_decay.assign(o->D, 1.0);
  o->limits.assign(o->D, stomp_joint_limit{0, 0, 0.0, 0.0});
  o->noiseless_costs.assign(o->N, 0.0);
  o->collision_free_slots.assign(o->R + 1, 0);
  o->constraints_satisfied_slots.assign(o->R + 1, 1);
  *out = o;
  return 0;
}

int stomp_oracle_destroy(void* h) { delete static_cast<Oracle*>(h); return 0; }

int stomp_oracle_set_robot(void* h, const stomp_segment* segments, int32_t num_segments, int32_t reference_segment,
                           const stomp_sphere* spheres, int32_t num_spheres, const stomp_joint_limit* limits) {
  Oracle& o = *static_cast<Oracle*>(h);
  o.segs.assign(segments, segments + num_segments);
  for (int s = 0; s < num_segments; ++s)
    if (o.segs[s].parent >= s) return fail("segments must be in DFS pre-order");
  o.ref_seg = reference_segment;
  o.spheres.assign(spheres, spheres + num_spheres);
  o.K = num_spheres;
  if (limits) o.limits.assign(limits, limits + o.D);
  o.cp_pos.assign(size_t(o.Nall) * o.K * 3, 0.0);
  o.cp_pot.assign(size_t(o.Nall) * o.K, 0.0);
  o.cp_vel_mag.assign(size_t(o.Nall) * o.K, 0.0);
  o.cp_coll.assign(size_t(o.Nall) * o.K, 0);
  o.cp_vox.assign(size_t(o.Nall) * o.K * 3, 0);
  o.state_coll.assign(o.Nall, 0);
  return 0;
}

int stomp_oracle_set_sdf(void* h, const void* voxels, int32_t nx, int32_t ny, int32_t nz, const double origin[3],
                         double resolution, int32_t voxel_dtype) {
  Oracle& o = *static_cast<Oracle*>(h);
  size_t cells = size_t(nx) * ny * nz;
  size_t bytes = cells * (voxel_dtype == STOMP_VOXEL_U8_SQ ? 1 : voxel_dtype == STOMP_VOXEL_U16_SQ ? 2 : 4);
  o.vox.assign(static_cast<const uint8_t*>(voxels), static_cast<const uint8_t*>(voxels) + bytes);
  o.nx = nx; o.ny = ny; o.nz = nz;
  o.vox_dtype = voxel_dtype;
  o.res = resolution;
  for (int i = 0; i < 3; ++i) o.origin[i] = origin[i];
  o.sqrt_table.resize(65536);
  for (int i = 0; i < 65536; ++i) o.sqrt_table[i] = std::sqrt(double(i)) * resolution; /* PropagationDistanceField sqrt_table_ */
  return 0;
}

int stomp_oracle_set_constraints(void* h, const stomp_orientation_constraint* c, int32_t n, double weight) {
  Oracle& o = *static_cast<Oracle*>(h);
  for (int i = 0; i < n; ++i)
    if (c[i].segment < 0 || c[i].segment >= int(o.segs.size())) return fail("constraint segment out of range");
  o.constraints.assign(c, c + n);
  o.constraint_cost_weight = weight;
  return 0;
}

int stomp_oracle_set_dynamics(void* h, const stomp_link_inertia* inertia, int32_t root, int32_t tip, const double* gravity,
                              double torque_cost_weight) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<int> path;
  for (int s = tip; s != root; s = o.segs[s].parent) {
    if (s < 0 || s >= int(o.segs.size())) return fail("chain tip is not below chain root");
    path.push_back(s);
  }
  o.chain.clear();
  int next_joint = 0;
  for (size_t k = path.size(); k-- > 0;) {
    const int i = path[k];
    if (o.segs[i].joint_type != STOMP_JOINT_FIXED) {
      if (o.segs[i].group_index != next_joint) return fail("the chain's movable joints must be the group joints in order");
      ++next_joint;
    }
    Oracle::ChainLink L;
    L.seg = i;
    L.m = inertia[i].mass;
    const double* c = inertia[i].com;
    const double* Ic = inertia[i].inertia;
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    for (int a = 0; a < 3; ++a) L.h[a] = L.m * c[a];
    /* inertia about the frame origin: Ic + m (|c|^2 1 - c c^T) */
    const double full[9] = {Ic[0], Ic[3], Ic[4], Ic[3], Ic[1], Ic[5], Ic[4], Ic[5], Ic[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) L.I[a * 3 + b] = full[a * 3 + b] + L.m * ((a == b ? cc : 0.0) - c[a] * c[b]);
    o.chain.push_back(L);
  }
  if (next_joint != o.D) return fail("the chain's movable joints must be the group joints in order");
  for (int a = 0; a < 3; ++a) o.gravity[a] = gravity[a];
  o.torque_cost_weight = torque_cost_weight;
  return 0;
}

/* joint torques [N][D] of the last execute (0 when the torque term is off) */
int stomp_oracle_last_torques(void* h, double* out) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::copy(o.last_torques.begin(), o.last_torques.end(), out);
  return 0;
}

int stomp_oracle_last_constraints_satisfied(void* h) { return static_cast<Oracle*>(h)->last_constraints_satisfied ? 1 : 0; }

int stomp_oracle_set_noise(void* h, const double* noise_stddev, const double* noise_decay) {
  Oracle& o = *static_cast<Oracle*>(h);
  o.noise_stddev.assign(noise_stddev, noise_stddev + o.D);
  o.noise_decay.assign(noise_decay, noise_decay + o.D);
  return 0;
}

int stomp_oracle_seed(void* h, uint64_t seed) {
  static_cast<Oracle*>(h)->rng.seed(uint32_t(seed));
  return 0;
}

/* StompOptimizer::initialize tail (src/stomp_optimizer.cpp:183-193) + the pre-loop part of
 * optimize() (src/stomp_optimizer.cpp:262-271): PI initialise, group trajectory padding, first FK. */
int stomp_oracle_set_problem(void* h, const double* start, const double* goal) {
  Oracle& o = *static_cast<Oracle*>(h);
  policy_set_to_min_control_cost(o, start, goal);
  if (!pi_initialize(o)) return fail(o.err);
  for (int d = 0; d < o.D; ++d) {
    for (int i = 0; i < o.fs; ++i) o.group_traj[d][i] = start[d];        /* src/stomp_trajectory.cpp:94-107 */
    for (int i = o.fe + 1; i < o.Nall; ++i) o.group_traj[d][i] = goal[d];
    for (int i = 0; i < o.N; ++i) o.group_traj[d][o.fs + i] = o.params_all[d][o.fs + i];
  }
  if (o.K > 0 && !o.vox.empty()) {
    o.iteration_ = 0;
    handle_joint_limits(o);
    perform_forward_kinematics(o);
  }
  return 0;
}

int stomp_oracle_set_parameters(void* h, const double* theta) {
  Oracle& o = *static_cast<Oracle*>(h);
  for (int d = 0; d < o.D; ++d)
    for (int i = 0; i < o.N; ++i) o.params_all[d][o.fs + i] = theta[size_t(d) * o.N + i];
  return 0;
}

int stomp_oracle_get_parameters(void* h, double* theta) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<Vec> p;
  policy_get_parameters(o, p);
  flatten(p, theta);
  return 0;
}

int stomp_oracle_update_parameters(void* h, const double* updates) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<Vec> u(o.D);
  for (int d = 0; d < o.D; ++d) u[d].assign(updates + size_t(d) * o.N, updates + size_t(d + 1) * o.N);
  policy_update_parameters(o, u);
  return 0;
}

int stomp_oracle_compute_control_costs(void* h, const double* parameters, const double* noise, int32_t n, double weight,
                                       double* control_costs) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<Vec> p(o.D), e(o.D), c(o.D);
  for (int r = 0; r < n; ++r) {
    for (int d = 0; d < o.D; ++d) {
      const double* pp = parameters + (size_t(r) * o.D + d) * o.N;
      const double* ee = noise + (size_t(r) * o.D + d) * o.N;
      p[d].assign(pp, pp + o.N);
      e[d].assign(ee, ee + o.N);
    }
    policy_compute_control_costs(o, p, e, weight, c);
    flatten(c, control_costs + size_t(r) * o.D * o.N);
  }
  return 0;
}

static std::vector<int> g_exec_satisfied;
int stomp_oracle_execute_constraints_satisfied(void*, int32_t* out, size_t count) {
  for (size_t i = 0; i < count && i < g_exec_satisfied.size(); ++i) out[i] = g_exec_satisfied[i];
  return 0;
}

int stomp_oracle_execute(void* h, const double* parameters, int32_t n, int32_t iteration_number, double* costs,
                         int32_t* collision_free) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<Vec> p(o.D);
  Vec c;
  for (int r = 0; r < n; ++r) {
    for (int d = 0; d < o.D; ++d) {
      const double* pp = parameters + (size_t(r) * o.D + d) * o.N;
      p[d].assign(pp, pp + o.N);
    }
    task_execute(o, p, c, iteration_number);
    std::copy(c.begin(), c.end(), costs + size_t(r) * o.N);
    if (collision_free) collision_free[r] = o.last_collision_free;
    if (r == 0) g_exec_satisfied.clear();
    g_exec_satisfied.push_back(o.last_constraints_satisfied ? 1 : 0);
  }
  return 0;
}

/* per-sphere records for trajectory points -1 .. N+1 of one rollout: debug[N+3][K] */
int stomp_oracle_execute_debug(void* h, const double* parameters, stomp_sphere_debug* debug, double* clipped) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<Vec> p(o.D);
  Vec c;
  for (int d = 0; d < o.D; ++d) p[d].assign(parameters + size_t(d) * o.N, parameters + size_t(d + 1) * o.N);
  task_execute(o, p, c, 1);
  for (int t = -1; t <= o.N + 1; ++t)
    for (int j = 0; j < o.K; ++j) {
      size_t src = size_t(o.fs + t) * o.K + j;
      stomp_sphere_debug& r = debug[size_t(t + 1) * o.K + j];
      for (int a = 0; a < 3; ++a) r.voxel[a] = o.cp_vox[src * 3 + a], r.position[a] = o.cp_pos[src * 3 + a];
      r.in_collision = o.cp_coll[src];
      r.potential = o.cp_pot[src];
      r.vel_mag = (t >= 0 && t < o.N) ? o.cp_vel_mag[src] : 0.0;
    }
  if (clipped)
    for (int d = 0; d < o.D; ++d)
      for (int i = 0; i < o.N; ++i) clipped[size_t(d) * o.N + i] = o.group_traj[d][o.fs + i];
  return 0;
}

int stomp_oracle_get_rollouts(void* h, const double* noise_stddev, const double* eps_injected, double* rollouts,
                              int32_t* num_generated) {
  Oracle& o = *static_cast<Oracle*>(h);
  pi_get_rollouts(o, noise_stddev, eps_injected);
  if (rollouts)
    for (int r = 0; r < o.num_rollouts_gen; ++r) flatten(o.rollouts[r].parameters, rollouts + size_t(r) * o.D * o.N);
  if (num_generated) *num_generated = o.num_rollouts_gen;
  return 0;
}

int stomp_oracle_set_rollout_costs(void* h, const double* costs, double control_cost_weight, double* totals) {
  pi_set_rollout_costs(*static_cast<Oracle*>(h), costs, control_cost_weight, totals);
  return 0;
}

int stomp_oracle_improve_policy(void* h, double* updates) {
  Oracle& o = *static_cast<Oracle*>(h);
  pi_improve_policy(o);
  if (updates) flatten(o.parameter_updates, updates);
  return 0;
}

int stomp_oracle_add_extra_rollouts(void* h, const double* costs) {
  Oracle& o = *static_cast<Oracle*>(h);
  std::vector<Vec> theta;
  policy_get_parameters(o, theta);
  pi_add_extra_rollout(o, theta, Vec(costs, costs + o.N));
  return 0;
}

int stomp_oracle_iterate(void* h, int32_t iteration_number, const double* eps_injected, double* noiseless_cost,
                         int32_t* noiseless_collision_free, int32_t* num_generated) {
  Oracle& o = *static_cast<Oracle*>(h);
  run_single_iteration(o, iteration_number, eps_injected);
  if (noiseless_cost) *noiseless_cost = o.last_cost;
  if (noiseless_collision_free) *noiseless_collision_free = o.last_collision_free;
  if (num_generated) *num_generated = o.num_rollouts_gen;
  return 0;
}

int stomp_oracle_get(void* h, int32_t field, void* out_, size_t bytes) {
  Oracle& o = *static_cast<Oracle*>(h);
  double* out = static_cast<double*>(out_);
  size_t RDN = size_t(o.R) * o.D * o.N, DN = size_t(o.D) * o.N, NN = size_t(o.N) * o.N;
  auto need = [&](size_t n) { return bytes >= n; };
  auto per_rollout = [&](std::vector<Vec> Rollout::*m) {
    for (int r = 0; r < o.R; ++r) flatten(o.rollouts[r].*m, out + size_t(r) * DN);
  };
  switch (field) {
    case STOMP_FIELD_THETA:
      if (!need(DN * 8)) return fail("buffer too small");
      return stomp_oracle_get_parameters(h, out);
    case STOMP_FIELD_NOISE: if (!need(RDN * 8)) return fail("buffer too small"); per_rollout(&Rollout::noise); return 0;
    case STOMP_FIELD_PARAMETERS: if (!need(RDN * 8)) return fail("buffer too small"); per_rollout(&Rollout::parameters); return 0;
    case STOMP_FIELD_NOISE_PROJECTED: if (!need(RDN * 8)) return fail("buffer too small"); per_rollout(&Rollout::noise_projected); return 0;
    case STOMP_FIELD_CONTROL_COSTS: if (!need(RDN * 8)) return fail("buffer too small"); per_rollout(&Rollout::control_costs); return 0;
    case STOMP_FIELD_CUMULATIVE_COSTS: if (!need(RDN * 8)) return fail("buffer too small"); per_rollout(&Rollout::cumulative_costs); return 0;
    case STOMP_FIELD_PROBABILITIES: if (!need(RDN * 8)) return fail("buffer too small"); per_rollout(&Rollout::probabilities); return 0;
    case STOMP_FIELD_STATE_COSTS:
      if (!need(size_t(o.R) * o.N * 8)) return fail("buffer too small");
      for (int r = 0; r < o.R; ++r) std::copy(o.rollouts[r].state_costs.begin(), o.rollouts[r].state_costs.end(), out + size_t(r) * o.N);
      return 0;
    case STOMP_FIELD_UPDATES: if (!need(DN * 8)) return fail("buffer too small"); flatten(o.parameter_updates, out); return 0;
    case STOMP_FIELD_NOISELESS_COSTS:
      if (!need(size_t(o.N) * 8)) return fail("buffer too small");
      std::copy(o.noiseless_costs.begin(), o.noiseless_costs.end(), out);
      return 0;
    case STOMP_FIELD_COLLISION_FREE:
      if (!need(size_t(o.R + 1) * 4)) return fail("buffer too small");
      std::copy(o.collision_free_slots.begin(), o.collision_free_slots.end(), static_cast<int32_t*>(out_));
      return 0;
    case STOMP_FIELD_CONSTRAINTS_SATISFIED:
      if (!need(size_t(o.R + 1) * 4)) return fail("buffer too small");
      std::copy(o.constraints_satisfied_slots.begin(), o.constraints_satisfied_slots.end(), static_cast<int32_t*>(out_));
      return 0;
    case STOMP_FIELD_ROLLOUT_TOTAL_COSTS:
      if (!need(size_t(o.R + 1) * 8)) return fail("buffer too small");
      for (int r = 0; r < o.R; ++r) out[r] = o.rollouts[r].getCost();
      out[o.R] = o.extra_rollouts[0].getCost();
      return 0;
    case STOMP_FIELD_INV_CONTROL_COST: if (!need(NN * 8)) return fail("buffer too small"); std::copy(o.Rinv.a.begin(), o.Rinv.a.end(), out); return 0;
    case STOMP_FIELD_NOISE_CHOLESKY: if (!need(NN * 8)) return fail("buffer too small"); std::copy(o.Lchol.a.begin(), o.Lchol.a.end(), out); return 0;
    case STOMP_FIELD_PROJECTION: if (!need(NN * 8)) return fail("buffer too small"); std::copy(o.Mproj.a.begin(), o.Mproj.a.end(), out); return 0;
    case STOMP_FIELD_QUAD_COST_INV: if (!need(NN * 8)) return fail("buffer too small"); std::copy(o.Qinv.a.begin(), o.Qinv.a.end(), out); return 0;
    case STOMP_FIELD_CONTROL_COST: if (!need(NN * 8)) return fail("buffer too small"); std::copy(o.Rfree.a.begin(), o.Rfree.a.end(), out); return 0;
    default: return fail("unknown field");
  }
}

/* ---- distance_field::PropagationDistanceField::addPointsToField, restated --------------------------------------------------
 * The reference rebuilds its field with this class (src/stomp_collision_space.cpp:187, distance_field_->addPointsToField); the
 * package is NOT in /root/reference (ROS `distance_field`, C-Turtle / Diamondback era, un-vendored), so this follows the
 * algorithm as that package publishes it: every voxel keeps {distance_square, closest obstacle cell, update direction};
 * obstacle cells enter bucket 0 with update direction (0,0,0); buckets are processed in order of squared distance; a voxel of
 * bucket 0 offers its closest point to all 26 neighbours, a voxel of a later bucket only to the 6-connected neighbours that do
 * not point against the direction it was itself updated from (dx*tdx >= 0, ...); a neighbour takes the offer when
 * |neighbour - closest|^2 is smaller than what it holds and at most max_distance_sq, and is queued in that bucket.  The
 * restricted neighbourhoods make this a propagation, not an exact transform: tests/test_oracle_kats.py counts where it
 * differs from the exact capped EDT the engine (and oracle/sdf_builder.py) compute.  occ / out are [nx][ny][nz]. */
int stomp_oracle_propagate_distance_field(int32_t nx, int32_t ny, int32_t nz, const uint8_t* occ, int32_t cap, int32_t* out_d2) {
  if (!occ || !out_d2 || nx < 1 || ny < 1 || nz < 1 || cap < 1) return fail("bad argument");
  const int max_sq = cap * cap;
  const size_t cells = size_t(nx) * ny * nz;
  struct Vox { int d2; int cp[3]; int dir; };
  std::vector<Vox> v(cells);
  for (size_t i = 0; i < cells; ++i) { v[i].d2 = max_sq; v[i].cp[0] = v[i].cp[1] = v[i].cp[2] = -1; v[i].dir = -1; }
  auto dirnum = [](int dx, int dy, int dz) { return (dx + 1) * 9 + (dy + 1) * 3 + dz + 1; };
  // neighbourhoods_[0][dir]: all 26 targets; neighbourhoods_[1][dir]: 6-connected targets not against dir
  std::vector<std::array<int, 3> > nb[2][27];
  for (int n = 0; n < 2; ++n)
    for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz)
      for (int tx = -1; tx <= 1; ++tx) for (int ty = -1; ty <= 1; ++ty) for (int tz = -1; tz <= 1; ++tz) {
        if (tx == 0 && ty == 0 && tz == 0) continue;
        if (n >= 1) {
          if (std::abs(tx) + std::abs(ty) + std::abs(tz) != 1) continue;
          if (dx * tx < 0 || dy * ty < 0 || dz * tz < 0) continue;
        }
        nb[n][dirnum(dx, dy, dz)].push_back({tx, ty, tz});
      }
  std::vector<std::vector<size_t> > buckets(size_t(max_sq) + 1);
  for (int x = 0; x < nx; ++x) for (int y = 0; y < ny; ++y) for (int z = 0; z < nz; ++z) {
    const size_t i = (size_t(x) * ny + y) * nz + z;
    if (!occ[i]) continue;
    v[i].d2 = 0; v[i].cp[0] = x; v[i].cp[1] = y; v[i].cp[2] = z; v[i].dir = dirnum(0, 0, 0);
    buckets[0].push_back(i);
  }
  for (int b = 0; b <= max_sq; ++b) {
    for (size_t q = 0; q < buckets[b].size(); ++q) {     // the bucket may grow while it is processed (offers at equal distance never win)
      const size_t i = buckets[b][q];
      const Vox cur = v[i];
      if (cur.dir < 0 || cur.dir > 26) continue;
      const int z = int(i % nz), y = int((i / nz) % ny), x = int(i / (size_t(nz) * ny));
      for (const auto& t : nb[b > 1 ? 1 : b][cur.dir]) {
        const int X = x + t[0], Y = y + t[1], Z = z + t[2];
        if (X < 0 || Y < 0 || Z < 0 || X >= nx || Y >= ny || Z >= nz) continue;
        const int ddx = X - cur.cp[0], ddy = Y - cur.cp[1], ddz = Z - cur.cp[2];
        const int nd = ddx * ddx + ddy * ddy + ddz * ddz;
        if (nd > max_sq) continue;
        Vox& nv = v[(size_t(X) * ny + Y) * nz + Z];
        if (nd < nv.d2) {
          nv.d2 = nd;
          nv.cp[0] = cur.cp[0]; nv.cp[1] = cur.cp[1]; nv.cp[2] = cur.cp[2];
          nv.dir = dirnum(t[0], t[1], t[2]);
          buckets[nd].push_back((size_t(X) * ny + Y) * nz + Z);
        }
      }
    }
    std::vector<size_t>().swap(buckets[b]);
  }
  for (size_t i = 0; i < cells; ++i) out_d2[i] = v[i].d2;
  return 0;
}

} /* extern "C" */
