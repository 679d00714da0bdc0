/*
 * oracle/ref_driver.cpp — C entry points around the reference's OWN PI^2 classes, compiled unmodified
 * from /root/reference/stomp_motion_planner/src/{policy_improvement,policy_improvement_loop,
 * covariant_trajectory_policy,stomp_cost}.cpp against the stand-in headers in oracle/ref_shim/.
 *
 * THIS IS TEST INFRASTRUCTURE (oracle/_ref/libstomp_ref_pi2.so, built by oracle/Makefile when
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
#include <stomp_motion_planner/policy_improvement_loop.h>
#include <stomp_motion_planner/covariant_trajectory_policy.h>
#include <stomp_motion_planner/stomp_cost.h>
#undef private
#undef protected

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
int stomp_ref_get(void* p, const char* field, double* out) {
  Ref* h = static_cast<Ref*>(p);
  PolicyImprovement& pi = h->loop.policy_improvement_;
  std::string f(field);
  const int N = h->N, D = h->D;
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
  if (f == "control_cost_matrix_all") { copy_mat_rowmajor(h->policy->control_costs_all_[0], out); return 0; }
  if (f == "parameters_all") {
    for (int d = 0; d < D; ++d) copy_vec(h->policy->parameters_all_[d], out + size_t(d) * (N + 12));
    return 0;
  }
  if (f == "parameter_updates") {
    for (int d = 0; d < D; ++d)
      for (int t = 0; t < N; ++t) out[size_t(d) * N + t] = pi.parameter_updates_[d](0, t);
    return 0;
  }
  if (f == "num_rollouts_gen") { out[0] = pi.num_rollouts_gen_; return 0; }
  if (f == "movement_dt") { out[0] = h->policy->movement_dt_; return 0; }
  return 2;
}

/* StompCost as StompOptimizer::initialize builds and scales it (src/stomp_optimizer.cpp:104-125) for
 * n_joints joints with per-joint cost multipliers joint_cost[j]; out = quad_cost_inv_ [n_joints][N][N]. */
int stomp_ref_quad_cost_inv(int num_vars_all, double discretization, const double* smoothness_costs /*[3]*/,
                            double ridge, int n_joints, const double* joint_cost, double* out) {
  StompTrajectory traj(num_vars_all, discretization);
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
