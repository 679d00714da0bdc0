/*
 * oracle/ref_driver.cpp — C entry points around the reference's OWN PI^2 classes, compiled unmodified
 * from /root/reference/stomp_motion_planner/src/{policy_improvement,policy_improvement_loop,
 * covariant_trajectory_policy,stomp_cost}.cpp against the stand-in headers in oracle/ref_shim/.
 *
 * THIS IS TEST INFRASTRUCTURE (oracle/_ref/libstomp_ref.so, built by oracle/Makefile when
 * /root/reference is present).  It exists to pin oracle/stomp_oracle.cpp — and through it the CUDA
 * engine — to outputs of the reference itself.  Nothing under stomp_motion_planner_icra2011_b200/ may
 * load it.  tests/golden/make_ref_golden.py runs it here and commits the vectors it produces, because
 * /root/reference does not exist on the GPU box.
 *
 * What runs from the reference, line for line: PolicyImprovementLoop::runSingleIteration, all of
 * PolicyImprovement (reuse sort, noise = L z, M eps, control costs, cumulative costs, probabilities,
 * updates), CovariantTrajectoryPolicy (differentiation matrices, R, min-control-cost trajectory,
 * control-cost folding, updateParameters), MultivariateGaussian, StompCost (joint-limit Q^-1).
 * What does not: the cost plugin (StompOptimizer::execute needs KDL, distance_field and the ROS
 * planning stack) — the Task below calls back into the test for state costs.
 */
#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <Eigen/Core>
#include <ros/ros.h>
#include <boost/shared_ptr.hpp>
#include <boost/random/variate_generator.hpp>
#include <boost/random/normal_distribution.hpp>
#include <boost/random/mersenne_twister.hpp>
#define private public   /* read Rollout / PolicyImprovement internals; the reference TUs are compiled without this */
#define protected public
#include <kdl/tree.hpp>
#include <ros_msgs_shim.h>
#include <stomp_motion_planner/policy_improvement_loop.h>
#include <stomp_motion_planner/covariant_trajectory_policy.h>
#include <stomp_motion_planner/stomp_cost.h>
#include <stomp_motion_planner/stomp_optimizer.h>
#include <stomp_motion_planner/stomp_collision_space.h>
#include <stomp_motion_planner/stomp_parameters.h>
#include <stomp_motion_planner/STOMPStatistics.h>
#include <planning_environment/monitors/collision_space_monitor.h>
#undef private
#undef protected

#include "../include/stomp_b200.h" /* plain-C table structs shared with the engine and the oracle */

#include <cstring>
#include <string>
#include <vector>

using namespace stomp_motion_planner;

extern "C" {
typedef int (*stomp_ref_execute_cb)(void* user, const double* parameters /*[D][N]*/, double* costs /*[N]*/,
                                    int iteration_number);
}

namespace {

class CallbackTask : public Task {
 public:
  CallbackTask(boost::shared_ptr<CovariantTrajectoryPolicy> policy, double control_cost_weight, int D, int N,
               stomp_ref_execute_cb cb, void* user)
      : policy_(policy), w_(control_cost_weight), D_(D), N_(N), cb_(cb), user_(user), buf_(size_t(D) * N),
        out_(size_t(N)) {}
  /* StompOptimizer::initialize(ros::NodeHandle&, int) does nothing (src/stomp_optimizer.cpp:1000-1004) */
  bool initialize(ros::NodeHandle&, int) { return true; }
  bool execute(std::vector<Eigen::VectorXd>& parameters, Eigen::VectorXd& costs, const int iteration_number) {
    for (int d = 0; d < D_; ++d)
      for (int t = 0; t < N_; ++t) buf_[size_t(d) * N_ + t] = parameters[d](t);
    int rc = cb_(user_, buf_.data(), out_.data(), iteration_number);
    for (int t = 0; t < N_; ++t) costs(t) = out_[t];
    return rc == 0;
  }
  bool getPolicy(boost::shared_ptr<Policy>& policy) { policy = policy_; return true; }
  bool setPolicy(const boost::shared_ptr<Policy>) { return true; }
  bool getControlCostWeight(double& w) { w = w_; return true; }

 private:
  boost::shared_ptr<CovariantTrajectoryPolicy> policy_;
  double w_;
  int D_, N_;
  stomp_ref_execute_cb cb_;
  void* user_;
  std::vector<double> buf_, out_;
};

struct Ref {
  int N, D, R;
  ros::NodeHandle nh;
  boost::shared_ptr<CovariantTrajectoryPolicy> policy;
  boost::shared_ptr<CallbackTask> task;
  PolicyImprovementLoop loop;
};

void copy_vec(const Eigen::VectorXd& v, double* out) {
  for (int i = 0; i < v.size(); ++i) out[i] = v(i);
}
void copy_mat_rowmajor(const Eigen::MatrixXd& m, double* out) {
  for (int i = 0; i < m.rows(); ++i)
    for (int j = 0; j < m.cols(); ++j) out[size_t(i) * m.cols() + j] = m(i, j);
}

}  // namespace

extern "C" {

/* Same order of construction as StompOptimizer::initialize (src/stomp_optimizer.cpp:181-191, 229-232):
 * policy->initialize(nh, N, D, duration, ridge, derivative_costs); setToMinControlCost(start, goal);
 * pi_loop.initialize(nh, task), which reads num_rollouts / num_reused_rollouts / num_time_steps /
 * noise_stddev / noise_decay / use_cumulative_costs from the parameter server. */
void* stomp_ref_create(int N, int D, int R, int R_reuse, double movement_duration, double ridge,
                       const double* derivative_costs, const double* noise_stddev, const double* noise_decay,
                       double control_cost_weight, int use_cumulative_costs, const double* start, const double* goal,
                       stomp_ref_execute_cb cb, void* user) {
  srand(1); /* MultivariateGaussian seeds mt19937 from rand() (multivariate_gaussian.h:84) */
  Ref* h = new Ref;
  h->N = N; h->D = D; h->R = R;
  h->nh = ros::NodeHandle("~");
  h->nh.set("num_rollouts", XmlRpc::XmlRpcValue(R));
  h->nh.set("num_reused_rollouts", XmlRpc::XmlRpcValue(R_reuse));
  h->nh.set("num_time_steps", XmlRpc::XmlRpcValue(N));
  h->nh.set("noise_stddev", XmlRpc::XmlRpcValue(std::vector<double>(noise_stddev, noise_stddev + D)));
  h->nh.set("noise_decay", XmlRpc::XmlRpcValue(std::vector<double>(noise_decay, noise_decay + D)));
  h->nh.set("use_cumulative_costs", XmlRpc::XmlRpcValue(bool(use_cumulative_costs != 0)));
  h->nh.set("write_to_file", XmlRpc::XmlRpcValue(false));

  h->policy.reset(new CovariantTrajectoryPolicy());
  std::vector<double> dc(derivative_costs, derivative_costs + 3);
  if (!h->policy->initialize(h->nh, N, D, movement_duration, ridge, dc)) { delete h; return 0; }
  Eigen::VectorXd s(D), g(D);
  for (int d = 0; d < D; ++d) { s(d) = start[d]; g(d) = goal[d]; }
  h->policy->setToMinControlCost(s, g);

  h->task.reset(new CallbackTask(h->policy, control_cost_weight, D, N, cb, user));
  if (!h->loop.initialize(h->nh, h->task)) { delete h; return 0; }
  return h;
}

void stomp_ref_destroy(void* p) { delete static_cast<Ref*>(p); }

int stomp_ref_run_single_iteration(void* p, int iteration_number) {
  return static_cast<Ref*>(p)->loop.runSingleIteration(iteration_number) ? 0 : 1;
}

int stomp_ref_set_parameters(void* p, const double* theta) {
  Ref* h = static_cast<Ref*>(p);
  std::vector<Eigen::VectorXd> v(h->D, Eigen::VectorXd::Zero(h->N));
  for (int d = 0; d < h->D; ++d)
    for (int t = 0; t < h->N; ++t) v[d](t) = theta[size_t(d) * h->N + t];
  return h->policy->setParameters(v) ? 0 : 1;
}

int stomp_ref_get_parameters(void* p, double* theta) {
  Ref* h = static_cast<Ref*>(p);
  std::vector<Eigen::VectorXd> v;
  if (!h->policy->getParameters(v)) return 1;
  for (int d = 0; d < h->D; ++d) copy_vec(v[d], theta + size_t(d) * h->N);
  return 0;
}

/* Policy::computeControlCosts(matrices, parameters, noise, weight, out) (covariant_trajectory_policy.cpp:228-255) */
int stomp_ref_compute_control_costs(void* p, const double* parameters, const double* noise, double weight,
                                    double* out) {
  Ref* h = static_cast<Ref*>(p);
  std::vector<Eigen::VectorXd> a(h->D, Eigen::VectorXd::Zero(h->N)), b(h->D, Eigen::VectorXd::Zero(h->N)),
      c(h->D, Eigen::VectorXd::Zero(h->N));
  for (int d = 0; d < h->D; ++d)
    for (int t = 0; t < h->N; ++t) {
      a[d](t) = parameters[size_t(d) * h->N + t];
      b[d](t) = noise[size_t(d) * h->N + t];
    }
  std::vector<Eigen::MatrixXd> mats;
  h->policy->getControlCosts(mats);
  if (!h->policy->computeControlCosts(mats, a, b, weight, c)) return 1;
  for (int d = 0; d < h->D; ++d) copy_vec(c[d], out + size_t(d) * h->N);
  return 0;
}

/* Read-out of internal state.  Per-rollout fields are [R][D][N] (state_costs [R][N], total [R]); the
 * "extra_" prefix addresses extra_rollouts_[0]; matrices are row-major [N][N] of dimension 0. */
static int pi_get(PolicyImprovement& pi, CovariantTrajectoryPolicy& policy, int N, int D, const char* field, double* out) {
  std::string f(field);
  std::vector<Rollout>* rs = &pi.rollouts_;
  if (f.compare(0, 6, "extra_") == 0) { rs = &pi.extra_rollouts_; f = f.substr(6); }
  const int R = int(rs->size());
  for (int r = 0; r < R; ++r) {
    Rollout& ro = (*rs)[r];
    if (f == "state_costs") { copy_vec(ro.state_costs_, out + size_t(r) * N); continue; }
    if (f == "total") { out[r] = ro.getCost(); continue; }
    std::vector<Eigen::VectorXd>* src = 0;
    if (f == "parameters") src = &ro.parameters_;
    else if (f == "noise") src = &ro.noise_;
    else if (f == "noise_projected") src = &ro.noise_projected_;
    else if (f == "control_costs") src = &ro.control_costs_;
    else if (f == "total_costs") src = &ro.total_costs_;
    else if (f == "cumulative_costs") src = &ro.cumulative_costs_;
    else if (f == "probabilities") src = &ro.probabilities_;
    if (!src) break;
    for (int d = 0; d < D; ++d) copy_vec((*src)[d], out + (size_t(r) * D + d) * N);
    if (r == R - 1) return 0;
  }
  if (f == "state_costs" || f == "total") return 0;
  if (f == "control_cost_matrix") { copy_mat_rowmajor(pi.control_costs_[0], out); return 0; }
  if (f == "inv_control_cost_matrix") { copy_mat_rowmajor(pi.inv_control_costs_[0], out); return 0; }
  if (f == "projection_matrix") { copy_mat_rowmajor(pi.projection_matrix_[0], out); return 0; }
  if (f == "covariance_cholesky") { copy_mat_rowmajor(pi.noise_generators_[0].covariance_cholesky_, out); return 0; }
  if (f == "control_cost_matrix_all") { copy_mat_rowmajor(policy.control_costs_all_[0], out); return 0; }
  if (f == "parameters_all") {
    for (int d = 0; d < D; ++d) copy_vec(policy.parameters_all_[d], out + size_t(d) * (N + 12));
    return 0;
  }
  if (f == "parameter_updates") {
    for (int d = 0; d < D; ++d)
      for (int t = 0; t < N; ++t) out[size_t(d) * N + t] = pi.parameter_updates_[d](0, t);
    return 0;
  }
  if (f == "num_rollouts_gen") { out[0] = pi.num_rollouts_gen_; return 0; }
  if (f == "movement_dt") { out[0] = policy.movement_dt_; return 0; }
  return 2;
}

int stomp_ref_get(void* p, const char* field, double* out) {
  Ref* h = static_cast<Ref*>(p);
  return pi_get(h->loop.policy_improvement_, *h->policy, h->N, h->D, field, out);
}

/* StompCost as StompOptimizer::initialize builds and scales it (src/stomp_optimizer.cpp:104-125) for
 * n_joints joints with per-joint cost multipliers joint_cost[j]; out = quad_cost_inv_ [n_joints][N][N]. */
int stomp_ref_quad_cost_inv(int num_vars_all, double discretization, const double* smoothness_costs /*[3]*/,
                            double ridge, int n_joints, const double* joint_cost, double* out) {
  StompRobotModel model;
  model.fk_solver_ = NULL;
  model.num_kdl_joints_ = 1;
  StompTrajectory traj(&model, num_vars_all, discretization);
  std::vector<StompCost> costs;
  costs.reserve(n_joints);
  double max_cost_scale = 0.0;
  for (int i = 0; i < n_joints; ++i) {
    std::vector<double> dc(3);
    for (int k = 0; k < 3; ++k) dc[k] = joint_cost[i] * smoothness_costs[k];
    costs.push_back(StompCost(traj, i, dc, ridge));
    double s = costs[i].getMaxQuadCostInvValue();
    if (max_cost_scale < s) max_cost_scale = s;
  }
  const int n = num_vars_all - 2 * (DIFF_RULE_LENGTH - 1);
  for (int i = 0; i < n_joints; ++i) {
    costs[i].scale(max_cost_scale);
    copy_mat_rowmajor(costs[i].getQuadraticCostInverse(), out + size_t(i) * n * n);
  }
  return 0;
}

} /* extern "C" */

/* =====================================================================================================
 * The cost-plugin half: the reference's StompOptimizer (src/stomp_optimizer.cpp, all 1213 lines), StompTrajectory,
 * StompCollisionPoint, both TreeFkSolverJointPosAxis solvers, OrientationConstraintEvaluator and StompParameters,
 * compiled unmodified.  src/stomp_robot_model.cpp and src/stomp_collision_space.cpp are compiled too, but their init()
 * functions need a URDF, a parameter server and a live planning_environment monitor; what they PRODUCE is exactly what the
 * engine's tables replace (include/stomp_b200.h: stomp_segment / stomp_sphere / stomp_joint_limit, the voxel grid), so the
 * driver fills the reference's own data structures (KDL::Tree, StompPlanningGroup, StompCollisionPoint list, distance field)
 * from those same tables, the way StompRobotModel::init (src/stomp_robot_model.cpp:95-190) and StompCollisionSpace::init
 * (src/stomp_collision_space.cpp:61-84) would.  The two generators that do run from those files are at the end of this file:
 * collision-point generation and collision-object rasterisation.
 * ===================================================================================================== */

namespace {

std::string seg_name(int i) { char b[32]; std::snprintf(b, sizeof b, "s%04d", i); return b; }

/* records the parameters of every Task::execute call, then runs the reference's own execute */
class SpyOptimizer : public StompOptimizer {
 public:
  SpyOptimizer(StompTrajectory* t, const StompRobotModel* m, const StompRobotModel::StompPlanningGroup* g,
               const StompParameters* p, const ros::Publisher& a, const ros::Publisher& b, const ros::Publisher& c,
               StompCollisionSpace* s, const motion_planning_msgs::Constraints& k)
      : StompOptimizer(t, m, g, p, a, b, c, s, k) {}
  bool execute(std::vector<Eigen::VectorXd>& parameters, Eigen::VectorXd& costs, const int iteration_number) {
    bool ok = StompOptimizer::execute(parameters, costs, iteration_number);
    std::vector<double> rec;
    for (size_t d = 0; d < parameters.size(); ++d)
      for (int t = 0; t < parameters[d].size(); ++t) rec.push_back(parameters[d](t));
    exec_iteration.push_back(iteration_number);
    exec_parameters.push_back(rec);
    std::vector<double> c(costs.size());
    for (int t = 0; t < costs.size(); ++t) c[t] = costs(t);
    exec_costs.push_back(c);
    exec_collision_free.push_back(last_trajectory_collision_free_ ? 1 : 0);
    exec_constraints_satisfied.push_back(last_trajectory_constraints_satisfied_ ? 1 : 0);
    return ok;
  }
  std::vector<int> exec_iteration, exec_collision_free, exec_constraints_satisfied;
  std::vector<std::vector<double> > exec_parameters, exec_costs;
};

struct RefOpt {
  int N, D, K, S;
  StompRobotModel model;
  StompParameters params;
  StompCollisionSpace space;
  distance_field::PropagationDistanceField* df;
  StompTrajectory* full;
  boost::shared_ptr<SpyOptimizer> opt;
  PolicyImprovementLoop* loop;
  ros::Publisher pub_a, pub_b, pub_stats;
  std::vector<uint8_t> voxels;
  std::vector<stomp_segment> segs;
  RefOpt() : df(NULL), full(NULL), loop(NULL) {}
  ~RefOpt() {
    if (opt) opt->resetSharedPtr();
    opt.reset();
    delete loop;
    delete full;
    space.distance_field_ = NULL;   /* ~StompCollisionSpace deletes its field; this one is ours */
    delete df;
    delete model.fk_solver_;
  }
};

}  // namespace

extern "C" {

void* stomp_ref_opt_create(const stomp_engine_desc* desc, const stomp_segment* segs, int32_t num_segments,
                           int32_t reference_segment, const stomp_sphere* spheres, int32_t num_spheres,
                           const stomp_joint_limit* limits, const void* voxels, int32_t nx, int32_t ny, int32_t nz,
                           const double* origin, double resolution, int32_t voxel_dtype, const double* noise_stddev,
                           const double* noise_decay, const double* start, const double* goal,
                           const stomp_orientation_constraint* constraints, int32_t num_constraints,
                           double constraint_cost_weight, int32_t max_iterations,
                           int32_t max_iterations_after_collision_free) {
  srand(1);
  RefOpt* h = new RefOpt;
  const int D = desc->num_dimensions, N = desc->num_time_steps;
  h->N = N; h->D = D; h->K = num_spheres; h->S = num_segments;
  h->segs.assign(segs, segs + num_segments);

  /* ---- KDL tree from the segment table (kdl_parser: Joint(name, origin, axis, type), Segment(name, joint, f_tip)) */
  StompRobotModel& m = h->model;
  m.fk_solver_ = NULL;
  m.num_kdl_joints_ = 0;
  m.max_radius_clearance_ = 0.0;
  m.kdl_tree_ = KDL::Tree(seg_name(0));
  for (int i = 1; i < num_segments; ++i) {
    const stomp_segment& g = segs[i];
    KDL::Vector pos(g.pos[0], g.pos[1], g.pos[2]), axis(g.axis[0], g.axis[1], g.axis[2]);
    KDL::Rotation rot(g.rot[0], g.rot[1], g.rot[2], g.rot[3], g.rot[4], g.rot[5], g.rot[6], g.rot[7], g.rot[8]);
    KDL::Joint joint = g.joint_type == STOMP_JOINT_REVOLUTE    ? KDL::Joint("j" + seg_name(i), pos, axis, KDL::Joint::RotAxis)
                       : g.joint_type == STOMP_JOINT_PRISMATIC ? KDL::Joint("j" + seg_name(i), pos, axis, KDL::Joint::TransAxis)
                                                               : KDL::Joint("j" + seg_name(i), KDL::Joint::None);
    if (!m.kdl_tree_.addSegment(KDL::Segment(seg_name(i), joint, KDL::Frame(rot, pos)), seg_name(g.parent))) {
      delete h;
      return 0;
    }
  }
  m.num_kdl_joints_ = m.kdl_tree_.getNrOfJoints();
  m.reference_frame_ = seg_name(reference_segment);
  m.fk_solver_ = new KDL::TreeFkSolverJointPosAxis(m.kdl_tree_, m.reference_frame_);
  m.kdl_number_to_urdf_name_.resize(m.num_kdl_joints_);
  for (int i = 1; i < num_segments; ++i)
    if (segs[i].joint_type != STOMP_JOINT_FIXED) {
      int q = m.kdl_tree_.getSegment(seg_name(i))->second.q_nr;
      m.kdl_number_to_urdf_name_[q] = "j" + seg_name(i);
      m.urdf_name_to_kdl_number_["j" + seg_name(i)] = q;
    }

  /* ---- planning group (src/stomp_robot_model.cpp:131-190) */
  StompRobotModel::StompPlanningGroup group;
  group.name_ = "group";
  group.num_joints_ = 0;
  group.stomp_joints_.resize(D);
  std::vector<bool> active_joints(m.num_kdl_joints_, false);
  for (int i = 1; i < num_segments; ++i) {
    int d = segs[i].group_index;
    if (d < 0) continue;
    StompRobotModel::StompJoint joint;
    KDL::SegmentMap::const_iterator it = m.kdl_tree_.getSegment(seg_name(i));
    joint.stomp_joint_index_ = d;
    joint.kdl_joint_index_ = it->second.q_nr;
    joint.kdl_joint_ = &(it->second.segment.getJoint());
    joint.link_name_ = seg_name(i);
    joint.joint_name_ = "j" + seg_name(i);
    joint.joint_update_limit_ = 0.1;
    joint.wrap_around_ = !limits[d].has_limits;
    joint.has_joint_limits_ = limits[d].has_limits != 0;
    joint.joint_limit_min_ = limits[d].min;
    joint.joint_limit_max_ = limits[d].max;
    group.stomp_joints_[d] = joint;
    group.num_joints_++;
    active_joints[joint.kdl_joint_index_] = true;
  }
  group.fk_solver_.reset(new KDL::TreeFkSolverJointPosAxisPartial(m.kdl_tree_, m.reference_frame_, active_joints));
  /* inverse dynamics: interface only (torque_cost_weight is 0; optimize() still calls it once per point for its
   * statistics message, src/stomp_optimizer.cpp:385-397) */
  group.id_solver_.reset(new KDL::ChainIdSolver_RNE(group.kdl_chain_, KDL::Vector(0, 0, -9.8)));
  /* collision points: parent joints by walking up the tree (StompRobotModel::getLinkInformation, :238-262) */
  for (int j = 0; j < num_spheres; ++j) {
    std::vector<int> parents;
    KDL::SegmentMap::const_iterator it = m.kdl_tree_.getSegment(seg_name(spheres[j].segment));
    while (it != m.kdl_tree_.getRootSegment()) {
      if (it->second.segment.getJoint().getType() != KDL::Joint::None) parents.push_back(it->second.q_nr);
      it = it->second.parent;
    }
    int segment_number = m.fk_solver_->segmentNameToIndex(seg_name(spheres[j].segment));
    group.collision_points_.push_back(StompCollisionPoint(parents, spheres[j].radius, spheres[j].clearance, segment_number,
                                                          KDL::Vector(spheres[j].pos[0], spheres[j].pos[1], spheres[j].pos[2])));
    if (m.max_radius_clearance_ < spheres[j].radius + spheres[j].clearance)
      m.max_radius_clearance_ = spheres[j].radius + spheres[j].clearance;
  }
  m.planning_groups_.insert(std::make_pair(group.name_, group));
  const StompRobotModel::StompPlanningGroup* pg = m.getPlanningGroup("group");

  /* ---- parameters (src/stomp_parameters.cpp:50-76 reads these from the parameter server) */
  ros::NodeHandle nh("~");
  nh.set("max_iterations", XmlRpc::XmlRpcValue(int(max_iterations)));
  nh.set("max_iterations_after_collision_free", XmlRpc::XmlRpcValue(int(max_iterations_after_collision_free)));
  nh.set("smoothness_cost_weight", XmlRpc::XmlRpcValue(desc->smoothness_cost_weight));
  nh.set("obstacle_cost_weight", XmlRpc::XmlRpcValue(desc->obstacle_cost_weight));
  nh.set("constraint_cost_weight", XmlRpc::XmlRpcValue(constraint_cost_weight));
  nh.set("torque_cost_weight", XmlRpc::XmlRpcValue(0.0));
  nh.set("smoothness_cost_velocity", XmlRpc::XmlRpcValue(desc->derivative_costs[0]));
  nh.set("smoothness_cost_acceleration", XmlRpc::XmlRpcValue(desc->derivative_costs[1]));
  nh.set("smoothness_cost_jerk", XmlRpc::XmlRpcValue(desc->derivative_costs[2]));
  nh.set("ridge_factor", XmlRpc::XmlRpcValue(desc->ridge_factor));
  nh.set("animate_path", XmlRpc::XmlRpcValue(false));
  nh.set("animate_endeffector", XmlRpc::XmlRpcValue(false));
  nh.set("animate_endeffector_segment", XmlRpc::XmlRpcValue(seg_name(num_segments - 1)));
  nh.set("use_chomp", XmlRpc::XmlRpcValue(false));
  nh.set("num_rollouts", XmlRpc::XmlRpcValue(int(desc->num_rollouts)));
  nh.set("num_reused_rollouts", XmlRpc::XmlRpcValue(int(desc->num_reused_rollouts)));
  nh.set("num_time_steps", XmlRpc::XmlRpcValue(N));
  nh.set("noise_stddev", XmlRpc::XmlRpcValue(std::vector<double>(noise_stddev, noise_stddev + D)));
  nh.set("noise_decay", XmlRpc::XmlRpcValue(std::vector<double>(noise_decay, noise_decay + D)));
  nh.set("use_cumulative_costs", XmlRpc::XmlRpcValue(bool(desc->use_cumulative_costs != 0)));
  nh.set("write_to_file", XmlRpc::XmlRpcValue(false));
  h->params.initFromNodeHandle();

  /* ---- collision space: the voxel grid the engine receives, behind the distance_field lookup */
  size_t vb = voxel_dtype == STOMP_VOXEL_U8_SQ ? 1 : (voxel_dtype == STOMP_VOXEL_U16_SQ ? 2 : 4);
  h->voxels.assign(static_cast<const uint8_t*>(voxels), static_cast<const uint8_t*>(voxels) + size_t(nx) * ny * nz * vb);
  h->df = new distance_field::PropagationDistanceField(h->voxels.data(), nx, ny, nz, origin, resolution, voxel_dtype);
  h->space.distance_field_ = h->df;
  h->space.field_bias_x_ = h->space.field_bias_y_ = h->space.field_bias_z_ = 0.0;   /* collision_space/field_bias_* default */
  h->space.resolution_ = resolution;
  h->space.reference_frame_ = m.reference_frame_;

  /* ---- full trajectory as StompPlannerNode::planKinematicPath builds it (src/stomp_planner_node.cpp:188-220):
   *      N+2 points, start state at 0 (all robot joints), goal at the last point, min-jerk in between */
  h->full = new StompTrajectory(&m, N + 2, desc->discretization);
  for (int i = 1; i < num_segments; ++i) {
    if (segs[i].joint_type == STOMP_JOINT_FIXED) continue;
    int q = m.kdl_tree_.getSegment(seg_name(i))->second.q_nr;
    int d = segs[i].group_index;
    (*h->full)(0, q) = d >= 0 ? start[d] : segs[i].fixed_value;
  }
  int goal_index = h->full->getNumPoints() - 1;
  h->full->getTrajectoryPoint(goal_index) = h->full->getTrajectoryPoint(0);
  for (int d = 0; d < D; ++d) (*h->full)(goal_index, pg->stomp_joints_[d].kdl_joint_index_) = goal[d];
  h->full->fillInMinJerk();

  motion_planning_msgs::Constraints cons;
  for (int i = 0; i < num_constraints; ++i) {
    motion_planning_msgs::OrientationConstraint oc;
    oc.link_name = seg_name(constraints[i].segment);
    oc.type = constraints[i].body_fixed ? int(motion_planning_msgs::OrientationConstraint::LINK_FRAME)
                                        : int(motion_planning_msgs::OrientationConstraint::HEADER_FRAME);
    oc.orientation.x = constraints[i].orientation[0]; oc.orientation.y = constraints[i].orientation[1];
    oc.orientation.z = constraints[i].orientation[2]; oc.orientation.w = constraints[i].orientation[3];
    oc.absolute_roll_tolerance = constraints[i].absolute_roll_tolerance;
    oc.absolute_pitch_tolerance = constraints[i].absolute_pitch_tolerance;
    oc.absolute_yaw_tolerance = constraints[i].absolute_yaw_tolerance;
    oc.weight = constraints[i].weight;
    cons.orientation_constraints.push_back(oc);
  }

  h->opt.reset(new SpyOptimizer(h->full, &m, pg, &h->params, h->pub_a, h->pub_b, h->pub_stats, &h->space, cons));
  boost::shared_ptr<StompOptimizer> base = h->opt;
  h->opt->setSharedPtr(base);
  return h;
}

void stomp_ref_opt_destroy(void* p) { delete static_cast<RefOpt*>(p); }

/* KDL::Chain + ChainIdSolver_RNE of the planning group (src/stomp_robot_model.cpp:181-185, where the chain is hard-coded to
 * the PR2's torso_lift_link -> r_gripper_tool_frame) for the path root -> tip of the table, with the segment inertias, and
 * the torque_cost_weight parameter. */
int stomp_ref_opt_set_dynamics(void* p, const stomp_link_inertia* inertia, int32_t root, int32_t tip, const double* gravity,
                               double torque_cost_weight) {
  RefOpt* h = static_cast<RefOpt*>(p);
  std::vector<int> path;
  for (int s = tip; s != root; s = h->segs[s].parent) {
    if (s < 0) return 1;
    path.push_back(s);
  }
  KDL::Chain chain;
  for (size_t k = path.size(); k-- > 0;) {
    const int i = path[k];
    const stomp_segment& g = h->segs[i];
    KDL::Vector pos(g.pos[0], g.pos[1], g.pos[2]), axis(g.axis[0], g.axis[1], g.axis[2]);
    KDL::Rotation rot(g.rot[0], g.rot[1], g.rot[2], g.rot[3], g.rot[4], g.rot[5], g.rot[6], g.rot[7], g.rot[8]);
    KDL::Joint joint = g.joint_type == STOMP_JOINT_REVOLUTE    ? KDL::Joint("j" + seg_name(i), pos, axis, KDL::Joint::RotAxis)
                       : g.joint_type == STOMP_JOINT_PRISMATIC ? KDL::Joint("j" + seg_name(i), pos, axis, KDL::Joint::TransAxis)
                                                               : KDL::Joint("j" + seg_name(i), KDL::Joint::None);
    KDL::RigidBodyInertia rbi(inertia[i].mass, KDL::Vector(inertia[i].com[0], inertia[i].com[1], inertia[i].com[2]),
                              inertia[i].inertia);
    chain.addSegment(KDL::Segment(seg_name(i), joint, KDL::Frame(rot, pos), rbi));
  }
  StompRobotModel::StompPlanningGroup& group = h->model.planning_groups_.find("group")->second;
  group.kdl_chain_ = chain;
  group.id_solver_.reset(new KDL::ChainIdSolver_RNE(group.kdl_chain_, KDL::Vector(gravity[0], gravity[1], gravity[2])));
  h->params.torque_cost_weight_ = torque_cost_weight;
  return int(chain.getNrOfJoints()) == h->D ? 0 : 2;
}

/* numbers the test checks before trusting anything else: [0] derived policy movement duration (int-truncated group
 * duration, src/stomp_optimizer.cpp:185), [1] num_vars_free, [2] num_vars_all, [3] free_vars_start, [4] #segments of the FK
 * solver, [5] 1 if the solver numbers segment i of the table as i (DFS pre-order), [6] #kdl joints */
int stomp_ref_opt_info(void* p, double* out) {
  RefOpt* h = static_cast<RefOpt*>(p);
  out[0] = h->opt->group_trajectory_.getDuration();
  out[1] = h->opt->num_vars_free_;
  out[2] = h->opt->num_vars_all_;
  out[3] = h->opt->free_vars_start_;
  out[4] = h->model.fk_solver_->getSegmentNames().size();
  bool same = true;
  for (int i = 0; i < h->S; ++i) same = same && h->model.fk_solver_->segmentNameToIndex(seg_name(i)) == i;
  out[5] = same ? 1 : 0;
  out[6] = h->model.num_kdl_joints_;
  return 0;
}

/* StompOptimizer::execute (src/stomp_optimizer.cpp:1063-1165) with iteration_ = iteration_number - 1, as inside optimize() */
int stomp_ref_opt_execute(void* p, const double* parameters, int32_t iteration_number, double* costs,
                          int32_t* collision_free, int32_t* constraints_satisfied) {
  RefOpt* h = static_cast<RefOpt*>(p);
  std::vector<Eigen::VectorXd> v(h->D, Eigen::VectorXd::Zero(h->N));
  for (int d = 0; d < h->D; ++d)
    for (int t = 0; t < h->N; ++t) v[d](t) = parameters[size_t(d) * h->N + t];
  Eigen::VectorXd c = Eigen::VectorXd::Zero(h->N);
  h->opt->iteration_ = iteration_number - 1;
  bool ok = h->opt->StompOptimizer::execute(v, c, iteration_number);
  for (int t = 0; t < h->N; ++t) costs[t] = c(t);
  if (collision_free) *collision_free = h->opt->last_trajectory_collision_free_ ? 1 : 0;
  if (constraints_satisfied) *constraints_satisfied = h->opt->last_trajectory_constraints_satisfied_ ? 1 : 0;
  return ok ? 0 : 1;
}

/* per-sphere state after the last execute, trajectory points -1 .. N+1: debug[N+3][K]; clipped[D][N] = group trajectory
 * after handleJointLimits.  voxel = the distance field's own cell rule applied to the reference's sphere position. */
int stomp_ref_opt_debug(void* p, stomp_sphere_debug* debug, double* clipped) {
  RefOpt* h = static_cast<RefOpt*>(p);
  StompOptimizer& o = *h->opt;
  for (int t = -1; t <= h->N + 1; ++t)
    for (int j = 0; j < h->K; ++j) {
      int i = o.free_vars_start_ + t;
      stomp_sphere_debug& r = debug[size_t(t + 1) * h->K + j];
      for (int a = 0; a < 3; ++a) {
        r.position[a] = o.collision_point_pos_[i][j](a);
        r.voxel[a] = h->df->getCellFromLocation(a, r.position[a]);
      }
      r.in_collision = o.point_is_in_collision_[i][j];
      r.potential = o.collision_point_potential_[i][j];
      r.vel_mag = (t >= 0 && t < h->N) ? o.collision_point_vel_mag_[i][j] : 0.0;
    }
  if (clipped)
    for (int d = 0; d < h->D; ++d)
      for (int t = 0; t < h->N; ++t) clipped[size_t(d) * h->N + t] = o.group_trajectory_(o.free_vars_start_ + t, d);
  return 0;
}

/* segment frames of trajectory point t (0-based free index) after the last execute: out[S][12] = rot row-major, pos */
int stomp_ref_opt_frames(void* p, int32_t t, double* out) {
  RefOpt* h = static_cast<RefOpt*>(p);
  const std::vector<KDL::Frame>& f = h->opt->segment_frames_[h->opt->free_vars_start_ + t];
  for (int s = 0; s < h->S; ++s) {  /* table order; the solver's own number of table segment s may differ */
    const KDL::Frame& fr = f[h->model.fk_solver_->segmentNameToIndex(seg_name(s))];
    for (int k = 0; k < 9; ++k) out[s * 12 + k] = fr.M.data[k];
    for (int k = 0; k < 3; ++k) out[s * 12 + 9 + k] = fr.p(k);
  }
  return 0;
}

/* The prologue of StompOptimizer::optimize (src/stomp_optimizer.cpp:261-271) with the PI loop owned by the handle, so
 * that its internals can be read between iterations. */
int stomp_ref_opt_begin(void* p) {
  RefOpt* h = static_cast<RefOpt*>(p);
  delete h->loop;
  h->loop = new PolicyImprovementLoop;
  ros::NodeHandle nh("~");
  boost::shared_ptr<Task> task = h->opt;
  if (!h->loop->initialize(nh, task)) return 1;
  StompOptimizer& o = *h->opt;
  o.iteration_ = 0;
  o.copyPolicyToGroupTrajectory();
  o.handleJointLimits();
  o.updateFullTrajectory();
  o.performForwardKinematics();
  return 0;
}

/* one pass of the loop body of optimize(): iteration_ = iteration_number - 1; pi_loop.runSingleIteration(iteration_ + 1) */
int stomp_ref_opt_iterate(void* p, int32_t iteration_number, double* noiseless_cost, int32_t* collision_free,
                          int32_t* constraints_satisfied) {
  RefOpt* h = static_cast<RefOpt*>(p);
  h->opt->iteration_ = iteration_number - 1;
  if (!h->loop->runSingleIteration(iteration_number)) return 1;
  *noiseless_cost = h->opt->last_trajectory_cost_;
  *collision_free = h->opt->last_trajectory_collision_free_ ? 1 : 0;
  *constraints_satisfied = h->opt->last_trajectory_constraints_satisfied_ ? 1 : 0;
  return 0;
}

/* same selector strings as stomp_ref_get, on the loop begun by stomp_ref_opt_begin */
int stomp_ref_opt_get(void* p, const char* field, double* out) {
  RefOpt* h = static_cast<RefOpt*>(p);
  std::string f(field);
  if (f == "exec_collision_free") {
    for (size_t i = 0; i < h->opt->exec_collision_free.size(); ++i) out[i] = h->opt->exec_collision_free[i];
    return 0;
  }
  if (f == "exec_count") { out[0] = h->opt->exec_costs.size(); return 0; }
  if (f == "exec_clear") {
    h->opt->exec_iteration.clear(); h->opt->exec_parameters.clear(); h->opt->exec_costs.clear();
    h->opt->exec_collision_free.clear(); h->opt->exec_constraints_satisfied.clear();
    return 0;
  }
  if (f == "exec_parameters") {
    for (size_t i = 0; i < h->opt->exec_parameters.size(); ++i)
      std::copy(h->opt->exec_parameters[i].begin(), h->opt->exec_parameters[i].end(), out + i * size_t(h->D) * h->N);
    return 0;
  }
  if (f == "exec_costs") {
    for (size_t i = 0; i < h->opt->exec_costs.size(); ++i)
      std::copy(h->opt->exec_costs[i].begin(), h->opt->exec_costs[i].end(), out + i * size_t(h->N));
    return 0;
  }
  if (f == "best_group_trajectory") {
    for (int d = 0; d < h->D; ++d)
      for (int t = 0; t < h->N; ++t) out[size_t(d) * h->N + t] = h->opt->group_trajectory_(h->opt->free_vars_start_ + t, d);
    return 0;
  }
  if (f == "theta") {
    std::vector<Eigen::VectorXd> v;
    h->opt->policy_->getParameters(v);
    for (int d = 0; d < h->D; ++d) copy_vec(v[d], out + size_t(d) * h->N);
    return 0;
  }
  if (f == "quad_cost_inv") { copy_mat_rowmajor(h->opt->joint_costs_[0].getQuadraticCostInverse(), out); return 0; }
  if (!h->loop) return 3;
  return pi_get(h->loop->policy_improvement_, *h->opt->policy_, h->N, h->D, field, out);
}

/* The reference's whole StompOptimizer::optimize() (src/stomp_optimizer.cpp:248-400).  stats[0..5] = success,
 * success_iteration, collision_success_iteration, best_cost, #iterations run (costs.size()), last_improvement_iteration;
 * costs[] = STOMPStatistics::costs (caller sizes it to max_iterations). */
int stomp_ref_opt_optimize(void* p, double* stats, double* costs) {
  RefOpt* h = static_cast<RefOpt*>(p);
  h->opt->optimize();
  boost::shared_ptr<STOMPStatistics> s = ros::Publisher::last<boost::shared_ptr<STOMPStatistics> >();
  if (!s) return 1;
  stats[0] = s->success ? 1 : 0;
  stats[1] = s->success_iteration;
  stats[2] = s->collision_success_iteration;
  stats[3] = s->best_cost;
  stats[4] = s->costs.size();
  stats[5] = h->opt->last_improvement_iteration_;
  for (size_t i = 0; i < s->costs.size(); ++i) costs[i] = s->costs[i];
  return 0;
}

} /* extern "C" */

/* =====================================================================================================
 * Generators from src/stomp_robot_model.cpp and src/stomp_collision_space.cpp, run unmodified on a mock
 * planning_environment (oracle/ref_shim/planning_environment/...): they produce the tables the engine consumes.
 * ===================================================================================================== */
extern "C" {

/* StompRobotModel::generateLinkCollisionPoints + generateAttachedObjectCollisionPoints + populatePlanningGroupCollisionPoints
 * (src/stomp_robot_model.cpp:352-496, 265-346).  links: the collision-checking links in getGroupLinkUnion() order with their
 * collision_links/<link>/{link_radius, link_clearance, link_extension} parameters (radius < 0: no link_radius parameter).
 * attached objects: owner segment, shape (0 sphere r | 1 box x y z | 2 cylinder r length), centre in the owner link's frame.
 * Writes the planning group's collision points (segment = table index) and returns their number. */
int stomp_ref_collision_points(const stomp_segment* segs, int32_t num_segments, int32_t reference_segment, int32_t D,
                               const int32_t* link_segment, const double* link_radius, const double* link_clearance,
                               const double* link_extension, int32_t num_links, double default_clearance,
                               const int32_t* att_segment, const int32_t* att_shape, const double* att_dims,
                               const double* att_position, int32_t num_attached, double attached_padding,
                               stomp_sphere* out, int32_t max_out) {
  StompRobotModel m;
  planning_environment::CollisionSpaceMonitor monitor;
  m.monitor_ = &monitor;
  m.kdl_tree_ = KDL::Tree(seg_name(0));
  for (int i = 1; i < num_segments; ++i) {
    const stomp_segment& g = segs[i];
    KDL::Vector pos(g.pos[0], g.pos[1], g.pos[2]), axis(g.axis[0], g.axis[1], g.axis[2]);
    KDL::Rotation rot(g.rot[0], g.rot[1], g.rot[2], g.rot[3], g.rot[4], g.rot[5], g.rot[6], g.rot[7], g.rot[8]);
    KDL::Joint joint = g.joint_type == STOMP_JOINT_REVOLUTE    ? KDL::Joint("j" + seg_name(i), pos, axis, KDL::Joint::RotAxis)
                       : g.joint_type == STOMP_JOINT_PRISMATIC ? KDL::Joint("j" + seg_name(i), pos, axis, KDL::Joint::TransAxis)
                                                               : KDL::Joint("j" + seg_name(i), KDL::Joint::None);
    if (!m.kdl_tree_.addSegment(KDL::Segment(seg_name(i), joint, KDL::Frame(rot, pos)), seg_name(g.parent))) return -1;
  }
  m.num_kdl_joints_ = m.kdl_tree_.getNrOfJoints();
  m.reference_frame_ = seg_name(reference_segment);
  m.fk_solver_ = new KDL::TreeFkSolverJointPosAxis(m.kdl_tree_, m.reference_frame_);
  m.collision_clearance_default_ = default_clearance;
  m.max_radius_clearance_ = 0.0;
  StompRobotModel::StompPlanningGroup group;
  group.name_ = "group";
  group.num_joints_ = D;
  group.stomp_joints_.resize(D);
  for (int i = 1; i < num_segments; ++i)
    if (segs[i].group_index >= 0) {
      group.stomp_joints_[segs[i].group_index].kdl_joint_index_ = m.kdl_tree_.getSegment(seg_name(i))->second.q_nr;
      group.stomp_joints_[segs[i].group_index].stomp_joint_index_ = segs[i].group_index;
    }
  m.planning_groups_.insert(std::make_pair(group.name_, group));

  ros::NodeHandle nh("~");
  std::vector<planning_models::KinematicModel::Link> owners(num_attached);
  std::vector<planning_models::KinematicModel::AttachedBody> bodies(num_attached);
  std::vector<shapes::Shape*> shapes;
  for (int i = 0; i < num_links; ++i) {
    const std::string name = seg_name(link_segment[i]);
    monitor.models.group_link_union.push_back(name);
    if (link_radius[i] >= 0.0) {
      nh.set("collision_links/" + name + "/link_radius", XmlRpc::XmlRpcValue(link_radius[i]));
      if (link_clearance[i] >= 0.0) nh.set("collision_links/" + name + "/link_clearance", XmlRpc::XmlRpcValue(link_clearance[i]));
      nh.set("collision_links/" + name + "/link_extension", XmlRpc::XmlRpcValue(link_extension[i]));
    }
  }
  m.generateLinkCollisionPoints();

  monitor.env.attached_padding = attached_padding;
  for (int i = 0; i < num_attached; ++i) {
    const double* d = att_dims + 3 * i;
    shapes::Shape* sh = att_shape[i] == 0   ? static_cast<shapes::Shape*>(new shapes::Sphere(d[0]))
                        : att_shape[i] == 1 ? static_cast<shapes::Shape*>(new shapes::Box(d[0], d[1], d[2]))
                                            : static_cast<shapes::Shape*>(new shapes::Cylinder(d[0], d[1]));
    shapes.push_back(sh);
    owners[i].name = seg_name(att_segment[i]);
    bodies[i].owner = &owners[i];
    bodies[i].shapes.push_back(sh);
    btTransform pose(btQuaternion(0, 0, 0, 1), btVector3(att_position[3 * i], att_position[3 * i + 1], att_position[3 * i + 2]));
    bodies[i].globalTrans.push_back(pose);   /* the mock tf listener is the identity: poses are given in the owner's frame */
    monitor.env.attached.push_back(&bodies[i]);
  }
  motion_planning_msgs::RobotState robot_state;
  m.generateAttachedObjectCollisionPoints(&robot_state);
  m.populatePlanningGroupCollisionPoints();

  const std::vector<StompCollisionPoint>& pts = m.getPlanningGroup("group")->collision_points_;
  std::map<int, int> number_to_table;
  for (int i = 0; i < num_segments; ++i) number_to_table[m.fk_solver_->segmentNameToIndex(seg_name(i))] = i;
  int n = 0;
  for (size_t i = 0; i < pts.size() && n < max_out; ++i, ++n) {
    out[n].segment = number_to_table[pts[i].getSegmentNumber()];
    out[n].radius = pts[i].getRadius();
    out[n].clearance = pts[i].getClearance();
    for (int k = 0; k < 3; ++k) out[n].pos[k] = pts[i].getPosition()(k);
  }
  for (size_t i = 0; i < shapes.size(); ++i) delete shapes[i];
  for (int i = 0; i < num_links; ++i) {   /* leave the shared parameter server clean for the next call */
    const std::string name = seg_name(link_segment[i]);
    nh.erase("collision_links/" + name + "/link_radius");
    nh.erase("collision_links/" + name + "/link_clearance");
    nh.erase("collision_links/" + name + "/link_extension");
  }
  delete m.fk_solver_;
  return int(pts.size());
}

/* StompCollisionSpace::addCollisionObjectsToPoints (src/stomp_collision_space.cpp:198-297) for boxes and cylinders, then the
 * distance field's addPointsToField binning: occupancy[nx][ny][nz] (nx = int(size / resolution)) of the cells the lattice
 * points fall into; returns the number of points generated (including those outside the grid). */
long long stomp_ref_collision_object_cells(const double* size, const double* origin, double resolution, const stomp_box* boxes,
                                           int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                           const double* map_points, int64_t num_map_points, uint8_t* occupancy, int32_t* dims) {
  StompCollisionSpace space;
  planning_environment::CollisionSpaceMonitor monitor;
  space.monitor_ = &monitor;
  space.resolution_ = resolution;
  distance_field::PropagationDistanceField df(size[0], size[1], size[2], resolution, origin[0], origin[1], origin[2], 0.0);
  space.distance_field_ = &df;
  collision_space::EnvironmentObjects::NamespaceObjects& no = monitor.env.env_objects.objects["objects"];
  std::vector<shapes::Shape*> shapes;
  for (int i = 0; i < num_boxes; ++i) {
    const stomp_box& b = boxes[i];
    shapes.push_back(new shapes::Box(b.dimensions[0], b.dimensions[1], b.dimensions[2]));
    no.shape.push_back(shapes.back());
    no.shapePose.push_back(btTransform(btQuaternion(b.orientation[0], b.orientation[1], b.orientation[2], b.orientation[3]),
                                       btVector3(b.position[0], b.position[1], b.position[2])));
  }
  for (int i = 0; i < num_cylinders; ++i) {
    const stomp_cylinder& c = cylinders[i];
    shapes.push_back(new shapes::Cylinder(c.radius, c.height));
    no.shape.push_back(shapes.back());
    no.shapePose.push_back(btTransform(btQuaternion(c.orientation[0], c.orientation[1], c.orientation[2], c.orientation[3]),
                                       btVector3(c.position[0], c.position[1], c.position[2])));
  }
  if (num_map_points > 0) {   /* the collision map arrives as the namespace "points": one (ignored) shape per point, pose = the point */
    collision_space::EnvironmentObjects::NamespaceObjects& np = monitor.env.env_objects.objects["points"];
    for (int64_t i = 0; i < num_map_points; ++i) {
      shapes.push_back(new shapes::Box(resolution, resolution, resolution));
      np.shape.push_back(shapes.back());
      np.shapePose.push_back(btTransform(btQuaternion(0, 0, 0, 1), btVector3(map_points[3 * i], map_points[3 * i + 1], map_points[3 * i + 2])));
    }
  }
  std::vector<btVector3> points;
  space.addCollisionObjectsToPoints(points);
  df.addPointsToField(points);
  for (int a = 0; a < 3; ++a) dims[a] = df.dim(a);
  if (occupancy) std::copy(df.occupied().begin(), df.occupied().end(), occupancy);
  for (size_t i = 0; i < shapes.size(); ++i) delete shapes[i];
  space.distance_field_ = NULL;
  return (long long)points.size();
}

} /* extern "C" */
