// host_math.hpp — once-per-request fp64 setup on the host: the smoothness matrix R of
// CovariantTrajectoryPolicy, its inverse / projection scaling, the banded Cholesky factor the
// device kernels solve with, StompCost's scaled inverse used by the joint-limit projection,
// and the folding of the KDL tree into the device joint table.
//
// Reference behaviour followed (paths relative to stomp_motion_planner/ in the reference):
//   src/covariant_trajectory_policy.cpp:150-226   dt, differentiation matrices, R_all, R, R^-1
//   src/policy_improvement.cpp:421-441            M = R^-1 with column p scaled by 1/(N*colmax_p)
//   src/stomp_cost.cpp:47-105, src/stomp_optimizer.cpp:105-125   quad_cost_inv_ and its scaling
//   src/treefksolverjointposaxis_partial.cpp:76-178              tree FK in a reference frame
//   include/stomp_motion_planner/stomp_utils.h:49-56            DIFF_RULES
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/stomp_b200.h"

namespace stomp_host {

static const int kRuleLen = STOMP_DIFF_RULE_LENGTH;
static const int kPad = STOMP_DIFF_RULE_LENGTH - 1;
static const double kDiffRules[STOMP_NUM_DIFF_RULES][STOMP_DIFF_RULE_LENGTH] = {
    {0, 0, -2 / 6.0, -3 / 6.0, 6 / 6.0, -1 / 6.0, 0},
    {0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0},
    {0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0}};

struct Dense {
  int n = 0;
  std::vector<double> a;
  Dense() {}
  explicit Dense(int n_) : n(n_), a(size_t(n_) * n_, 0.0) {}
  double& operator()(int i, int j) { return a[size_t(i) * n + j]; }
  double operator()(int i, int j) const { return a[size_t(i) * n + j]; }
};

// sum_k weight_k * (A_k^T A_k) for 7-tap stencil matrices A_k of size n (out-of-range taps dropped),
// scale_k multiplies the stencil of rule k.  Only the band is touched.
inline Dense stencil_gram(int n, const double weights[3], const double scales[3], double ridge) {
  Dense G(n);
  for (int k = 0; k < STOMP_NUM_DIFF_RULES; ++k) {
    if (weights[k] == 0.0) continue;
    for (int r = 0; r < n; ++r)
      for (int j1 = -3; j1 <= 3; ++j1) {
        int c1 = r + j1;
        if (c1 < 0 || c1 >= n) continue;
        double a1 = scales[k] * kDiffRules[k][j1 + 3];
        if (a1 == 0.0) continue;
        for (int j2 = -3; j2 <= 3; ++j2) {
          int c2 = r + j2;
          if (c2 < 0 || c2 >= n) continue;
          double a2 = scales[k] * kDiffRules[k][j2 + 3];
          G(c1, c2) += weights[k] * (a1 * a2);
        }
      }
  }
  for (int i = 0; i < n; ++i) G(i, i) += ridge;
  return G;
}

inline Dense free_block(const Dense& A, int n_free) {
  Dense B(n_free);
  for (int i = 0; i < n_free; ++i)
    for (int j = 0; j < n_free; ++j) B(i, j) = A(kPad + i, kPad + j);
  return B;
}

// LU inverse with partial pivoting (Eigen 2 MatrixXd::inverse() is LU based).
inline bool invert(const Dense& A, Dense& out) {
  int n = A.n;
  Dense W = A;
  out = Dense(n);
  for (int i = 0; i < n; ++i) out(i, i) = 1.0;
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

inline bool cholesky_lower(const Dense& A, Dense& L) {
  int n = A.n;
  L = Dense(n);
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

// Banded lower Cholesky R = C C^T.  band[i*(hb+1)+k] = C(i, i-k) for k=0..hb; band[...+0] holds the
// diagonal, inv_diag[i] = 1/C(i,i).
struct BandChol {
  int n = 0, hb = 0;
  std::vector<double> band, inv_diag;
};

inline bool band_cholesky(const Dense& R, BandChol& bc) {
  int n = R.n, hb = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (R(i, j) != 0.0) hb = std::max(hb, i - j);
  bc.n = n;
  bc.hb = hb;
  bc.band.assign(size_t(n) * (hb + 1), 0.0);
  bc.inv_diag.assign(n, 0.0);
  auto C = [&](int i, int j) -> double& { return bc.band[size_t(i) * (hb + 1) + (i - j)]; };
  for (int j = 0; j < n; ++j) {
    double s = R(j, j);
    for (int k = std::max(0, j - hb); k < j; ++k) s -= C(j, k) * C(j, k);
    if (s <= 0.0) return false;
    C(j, j) = std::sqrt(s);
    bc.inv_diag[j] = 1.0 / C(j, j);
    for (int i = j + 1; i <= std::min(n - 1, j + hb); ++i) {
      double t = R(i, j);
      for (int k = std::max(0, i - hb); k < j; ++k) t -= C(i, k) * C(j, k);
      C(i, j) = t * bc.inv_diag[j];
    }
  }
  return true;
}

struct PolicyMatrices {
  int N = 0, Nall = 0;
  double dt = 0.0;
  Dense R_all, R, Rinv, Qinv;
  std::vector<double> proj_scale;  // s_p = 1/(N * max_p2 Rinv(p2,p))
  BandChol chol;
  double stencil_scale[3];         // 1/dt^(k+1)
};

inline bool build_policy_matrices(const stomp_engine_desc& d, PolicyMatrices& pm, std::string& err) {
  pm.N = d.num_time_steps;
  pm.Nall = pm.N + 2 * kPad;
  pm.dt = d.movement_duration / (pm.N + 1);
  double mult = 1.0;
  for (int k = 0; k < 3; ++k) {
    mult /= pm.dt;
    pm.stencil_scale[k] = mult;
  }
  pm.R_all = stencil_gram(pm.Nall, d.derivative_costs, pm.stencil_scale, d.ridge_factor);
  pm.R = free_block(pm.R_all, pm.N);
  if (!invert(pm.R, pm.Rinv)) { err = "control cost matrix is singular"; return false; }
  if (!band_cholesky(pm.R, pm.chol)) { err = "control cost matrix is not positive definite"; return false; }
  pm.proj_scale.resize(pm.N);
  for (int p = 0; p < pm.N; ++p) {
    double column_max = pm.Rinv(0, p);
    for (int p2 = 1; p2 < pm.N; ++p2)
      if (pm.Rinv(p2, p) > column_max) column_max = pm.Rinv(p2, p);
    pm.proj_scale[p] = 1.0 / (pm.N * column_max);
  }
  // StompCost: unit-free stencils, weights derivative_costs[k] * discretization^(k+1), + ridge; inverse scaled
  // by the largest coefficient (all joint_costs are 1.0 so every joint shares the matrix).
  double w[3], ones[3] = {1.0, 1.0, 1.0}, m2 = 1.0;
  for (int k = 0; k < 3; ++k) {
    m2 *= d.discretization;
    w[k] = d.derivative_costs[k] * m2;
  }
  Dense Qall = stencil_gram(pm.Nall, w, ones, d.ridge_factor);
  Dense Q = free_block(Qall, pm.N);
  if (!invert(Q, pm.Qinv)) { err = "quad cost matrix is singular"; return false; }
  double mx = pm.Qinv.a[0];
  for (double v : pm.Qinv.a) mx = std::max(mx, v);
  double inv_scale = 1.0 / mx;
  for (double& v : pm.Qinv.a) v *= inv_scale;
  return true;
}

// setToMinControlCost for one dimension: theta = -0.5 * Rinv * lin, lin = 2*(start^T R_all[0:6, free] +
// goal^T R_all[fe+1:, free])  (src/covariant_trajectory_policy.cpp:102-148).
inline void min_control_cost(const PolicyMatrices& pm, double start, double goal, double* theta) {
  int N = pm.N, fs = kPad, fe = kPad + N - 1;
  std::vector<double> lin(N);
  for (int j = 0; j < N; ++j) {
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < kPad; ++i) s += start * pm.R_all(i, fs + j);
    for (int i = 0; i < kPad; ++i) s2 += goal * pm.R_all(fe + 1 + i, fs + j);
    lin[j] = 2.0 * (s + s2);
  }
  for (int k = 0; k < N; ++k) {
    double s = 0.0;
    const double* row = &pm.Rinv.a[size_t(k) * N];
    for (int j = 0; j < N; ++j) s += row[j] * lin[j];
    theta[k] = -0.5 * s;
  }
}

// ---- robot: fold the KDL-style tree into the device joint table --------------------------------
struct Frame {
  double R[9], p[3];
};
inline Frame frame_identity() {
  Frame f;
  std::memset(&f, 0, sizeof(f));
  f.R[0] = f.R[4] = f.R[8] = 1.0;
  return f;
}
inline Frame mul(const Frame& a, const Frame& b) {
  Frame c;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      c.R[i * 3 + j] = a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j] + a.R[i * 3 + 2] * b.R[6 + j];
    c.p[i] = a.R[i * 3] * b.p[0] + a.R[i * 3 + 1] * b.p[1] + a.R[i * 3 + 2] * b.p[2] + a.p[i];
  }
  return c;
}
inline Frame inverse(const Frame& a) {
  Frame c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) c.p[i] = -(c.R[i * 3] * a.p[0] + c.R[i * 3 + 1] * a.p[1] + c.R[i * 3 + 2] * a.p[2]);
  return c;
}
inline void rotate(const double R[9], const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}
inline void rodrigues(const double v[3], double angle, double R[9]) {
  double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
  R[0] = ct + vt * v[0] * v[0];
  R[1] = -v[2] * st + vt * v[0] * v[1];
  R[2] = v[1] * st + vt * v[0] * v[2];
  R[3] = v[2] * st + vt * v[0] * v[1];
  R[4] = ct + vt * v[1] * v[1];
  R[5] = -v[0] * st + vt * v[1] * v[2];
  R[6] = -v[1] * st + vt * v[0] * v[2];
  R[7] = v[0] * st + vt * v[1] * v[2];
  R[8] = ct + vt * v[2] * v[2];
}
inline Frame segment_pose(const stomp_segment& s, double q) {
  Frame f;
  if (s.joint_type == STOMP_JOINT_REVOLUTE) {
    double Rq[9];
    rodrigues(s.axis, q, Rq);
    Frame a, b;
    std::memcpy(a.R, Rq, sizeof(Rq));
    a.p[0] = a.p[1] = a.p[2] = 0.0;
    std::memcpy(b.R, s.rot, sizeof(b.R));
    b.p[0] = b.p[1] = b.p[2] = 0.0;
    f = mul(a, b);
    for (int i = 0; i < 3; ++i) f.p[i] = s.pos[i];
  } else {
    std::memcpy(f.R, s.rot, sizeof(f.R));
    for (int i = 0; i < 3; ++i) f.p[i] = s.pos[i] + (s.joint_type == STOMP_JOINT_PRISMATIC ? q * s.axis[i] : 0.0);
  }
  return f;
}

// One node of the device joint table = one group joint (or a static anchor that carries spheres).
// frame(node) = frame(parent node) * pose(q), pose(q).R = A0 + cos(q) A1 + sin(q) A2 (revolute) and
// pose(q).p = p + q * ax (prismatic).  All fixed transforms between consecutive group joints (and the
// reference-frame change for root nodes) are folded into A0..A2 / p / ax here, in fp64.
struct HostNode {
  int parent;      // node index, -1: the node's frame is pose(q) itself (static prefix folded in)
  int type;        // stomp_joint_type
  int q_index;     // group joint index, -1 for fixed
  int save_slot;   // >=0: this frame is needed later by a non-adjacent child -> stored in a local slot
  int load_slot;   // >=0: parent frame comes from this slot instead of the previous node
  int sphere_begin, sphere_end;  // range in the node-sorted sphere table
  double A0[9], A1[9], A2[9], p[3], ax[3];
};

struct HostSphere {
  int node;
  int original_index;
  double pos[3], radius, clearance, inv_clearance, weight;
};

struct FoldedRobot {
  std::vector<HostNode> nodes;
  std::vector<HostSphere> spheres;
  int num_slots = 0;
  // per input segment: the node its frame rides on (-1: static) and the frame relative to that node
  // (static segments: the constant frame in the reference frame)
  std::vector<int> seg_node;
  std::vector<Frame> seg_rel;
};

inline bool fold_robot(const stomp_segment* segs, int S, int ref_seg, const stomp_sphere* sph, int K, int D,
                       FoldedRobot& out, std::string& err) {
  if (ref_seg < 0 || ref_seg >= S) { err = "reference segment out of range"; return false; }
  std::vector<int> dyn(S, 0);  // 1 if a group joint lies on the path root..segment (inclusive)
  for (int s = 0; s < S; ++s) {
    if (segs[s].parent >= s) { err = "segments must be in DFS pre-order (parent index < own index)"; return false; }
    if (segs[s].group_index >= D) { err = "group_index out of range"; return false; }
    bool own = segs[s].group_index >= 0 && segs[s].joint_type != STOMP_JOINT_FIXED;
    dyn[s] = own || (segs[s].parent >= 0 && dyn[segs[s].parent]);
  }
  if (dyn[ref_seg]) { err = "the reference frame segment must not be moved by a group joint"; return false; }
  // static world frames, then into the reference frame
  std::vector<Frame> stat(S, frame_identity());
  for (int s = 0; s < S; ++s)
    if (!dyn[s]) stat[s] = mul(segs[s].parent >= 0 ? stat[segs[s].parent] : frame_identity(), segment_pose(segs[s], segs[s].fixed_value));
  Frame inv_ref = inverse(stat[ref_seg]);
  for (int s = 0; s < S; ++s)
    if (!dyn[s]) stat[s] = mul(inv_ref, stat[s]);
  // node_of[s]: node whose frame the segment is rigidly attached to; rel[s]: segment frame relative to it
  std::vector<int> node_of(S, -1);
  std::vector<Frame> rel(S, frame_identity());
  out.nodes.clear();
  for (int s = 0; s < S; ++s) {
    bool own = segs[s].group_index >= 0 && segs[s].joint_type != STOMP_JOINT_FIXED;
    if (!dyn[s]) continue;
    int par = segs[s].parent;
    if (!own) {  // rigidly attached to the parent's node
      node_of[s] = node_of[par];
      rel[s] = mul(rel[par], segment_pose(segs[s], segs[s].fixed_value));
      continue;
    }
    // new node.  Prefix = frame the joint is mounted on, relative to the parent node (or static, ref-framed)
    HostNode n;
    std::memset(&n, 0, sizeof(n));
    Frame prefix;
    if (par >= 0 && dyn[par]) { n.parent = node_of[par]; prefix = rel[par]; }
    else { n.parent = -1; prefix = par >= 0 ? stat[par] : inv_ref; }
    n.type = segs[s].joint_type;
    n.q_index = segs[s].group_index;
    n.save_slot = n.load_slot = -1;
    // pose(q) = Frame(Rot(a,q)*Rj, pj);  prefix*pose = Frame(Rot(P a, q) * P Rj, P pj + pp)
    double a[3], PR[9];
    rotate(prefix.R, segs[s].axis, a);
    Frame pr, rj;
    std::memcpy(pr.R, prefix.R, sizeof(pr.R));
    pr.p[0] = pr.p[1] = pr.p[2] = 0;
    std::memcpy(rj.R, segs[s].rot, sizeof(rj.R));
    rj.p[0] = rj.p[1] = rj.p[2] = 0;
    Frame prj = mul(pr, rj);
    std::memcpy(PR, prj.R, sizeof(PR));
    double pj[3];
    rotate(prefix.R, segs[s].pos, pj);
    for (int i = 0; i < 3; ++i) n.p[i] = pj[i] + prefix.p[i], n.ax[i] = a[i];
    if (n.type == STOMP_JOINT_REVOLUTE) {
      // Rot(a,q) PR = a a^T PR + cos q (PR - a a^T PR) + sin q [a]x PR
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double aaT = 0.0;
          for (int k = 0; k < 3; ++k) aaT += a[i] * a[k] * PR[k * 3 + j];
          n.A0[i * 3 + j] = aaT;
          n.A1[i * 3 + j] = PR[i * 3 + j] - aaT;
        }
      for (int j = 0; j < 3; ++j) {
        n.A2[0 * 3 + j] = a[1] * PR[2 * 3 + j] - a[2] * PR[1 * 3 + j];
        n.A2[1 * 3 + j] = a[2] * PR[0 * 3 + j] - a[0] * PR[2 * 3 + j];
        n.A2[2 * 3 + j] = a[0] * PR[1 * 3 + j] - a[1] * PR[0 * 3 + j];
      }
    } else {
      std::memcpy(n.A0, PR, sizeof(PR));
    }
    node_of[s] = int(out.nodes.size());
    rel[s] = frame_identity();
    out.nodes.push_back(n);
  }
  out.seg_node.assign(S, -1);
  out.seg_rel.assign(S, frame_identity());
  for (int s = 0; s < S; ++s) {
    out.seg_node[s] = dyn[s] ? node_of[s] : -1;
    out.seg_rel[s] = dyn[s] ? rel[s] : stat[s];
  }
  // spheres on static segments get an anchor node each distinct segment (fixed type, constant frame)
  std::vector<int> static_anchor(S, -1);
  out.spheres.clear();
  std::vector<HostSphere> tmp;
  for (int j = 0; j < K; ++j) {
    int s = sph[j].segment;
    if (s < 0 || s >= S) { err = "sphere segment out of range"; return false; }
    HostSphere hs;
    hs.original_index = j;
    hs.radius = sph[j].radius;
    hs.clearance = sph[j].clearance;
    hs.inv_clearance = 1.0 / sph[j].clearance;  // StompCollisionPoint ctor, src/stomp_collision_point.cpp:50
    hs.weight = double(K - j);                  // cumulative-over-j sum == sum_j (K-j) c_j, stomp_optimizer.cpp:1098-1105
    if (dyn[s]) {
      hs.node = node_of[s];
      double q[3];
      rotate(rel[s].R, sph[j].pos, q);
      for (int i = 0; i < 3; ++i) hs.pos[i] = q[i] + rel[s].p[i];
    } else {
      if (static_anchor[s] < 0) {
        HostNode n;
        std::memset(&n, 0, sizeof(n));
        n.parent = -1;
        n.type = STOMP_JOINT_FIXED;
        n.q_index = -1;
        n.save_slot = n.load_slot = -1;
        std::memcpy(n.A0, stat[s].R, sizeof(n.A0));
        for (int i = 0; i < 3; ++i) n.p[i] = stat[s].p[i];
        static_anchor[s] = int(out.nodes.size());
        out.nodes.push_back(n);
      }
      hs.node = static_anchor[s];
      for (int i = 0; i < 3; ++i) hs.pos[i] = sph[j].pos[i];
    }
    tmp.push_back(hs);
  }
  // sort spheres by node (stable: keeps the original order inside a node)
  int nn = int(out.nodes.size());
  for (int n = 0; n < nn; ++n) {
    out.nodes[n].sphere_begin = int(out.spheres.size());
    for (const HostSphere& hs : tmp)
      if (hs.node == n) out.spheres.push_back(hs);
    out.nodes[n].sphere_end = int(out.spheres.size());
  }
  // frame slots: a node whose parent is neither -1 nor the previous node loads the parent from a slot
  out.num_slots = 0;
  for (int n = 0; n < nn; ++n) {
    int par = out.nodes[n].parent;
    if (par >= 0 && par != n - 1) {
      if (out.nodes[par].save_slot < 0) out.nodes[par].save_slot = out.num_slots++;
      out.nodes[n].load_slot = out.nodes[par].save_slot;
    }
  }
  return true;
}

}  // namespace stomp_host
