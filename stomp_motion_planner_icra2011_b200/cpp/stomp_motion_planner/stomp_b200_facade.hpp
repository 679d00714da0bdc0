// stomp_b200_facade.hpp — host C++ classes with the reference's names and call signatures for the
// per-iteration rollout loop, implemented on top of the C ABI (include/stomp_b200.h).
//
// What maps to what (reference paths relative to stomp_motion_planner/):
//   Policy, Task                         include/stomp_motion_planner/policy.h:47-134, task.h:49-93
//   CovariantTrajectoryPolicy            include/stomp_motion_planner/covariant_trajectory_policy.h:48-283
//     (alias CovariantMovementPrimitive: the name BASELINE.json uses, from the later `stomp` package)
//   PolicyImprovement                    include/stomp_motion_planner/policy_improvement.h:65-126
//   PolicyImprovementLoop                include/stomp_motion_planner/policy_improvement_loop.h:53-62
//   StompOptimizer                       include/stomp_motion_planner/stomp_optimizer.h:63-114
//
// Differences that are forced by the missing dependencies (no ROS / Eigen 2 / Boost / KDL here):
//   Eigen::VectorXd -> std::vector<double>; Eigen::MatrixXd -> stomp_motion_planner::MatrixXd (row-major);
//   boost::shared_ptr -> std::shared_ptr; ros::NodeHandle parameters -> StompParameters (plain struct with the
//   reference's parameter names and defaults).  Every method returns bool like the reference; the C ABI's error
//   string is available from lastError().
//
// The cost-function plugin surface is Task::execute.  PolicyImprovementLoop::runSingleIteration runs the whole
// iteration on the GPU (one stomp_engine_iterate call) when the task is the built-in StompOptimizer, and falls
// back to the reference's exact call sequence (getRollouts -> task->execute per rollout -> setRolloutCosts ->
// improvePolicy -> updateParameters -> execute(noise-less) -> addExtraRollouts) for any other Task, with the
// PI^2 arithmetic still on the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "stomp_b200.h"

namespace stomp_motion_planner {

typedef std::vector<double> VectorXd;

struct MatrixXd {
  int rows_ = 0, cols_ = 0;
  std::vector<double> data;
  MatrixXd() {}
  MatrixXd(int r, int c) : rows_(r), cols_(c), data(size_t(r) * c, 0.0) {}
  static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  double& operator()(int i, int j) { return data[size_t(i) * cols_ + j]; }
  double operator()(int i, int j) const { return data[size_t(i) * cols_ + j]; }
};

inline const char* lastError() { return stomp_engine_last_error(); }

// ---- parameters: config/params.yaml + StompParameters defaults (src/stomp_parameters.cpp:53-75) ---------
struct StompParameters {
  double trajectory_duration = 5.0, trajectory_discretization = 0.05;
  int max_iterations = 500, max_iterations_after_collision_free = 500;
  double smoothness_cost_velocity = 0.0, smoothness_cost_acceleration = 1.0, smoothness_cost_jerk = 0.0;
  double smoothness_cost_weight = 0.1, obstacle_cost_weight = 1.0, ridge_factor = 0.0;
  double torque_cost_weight = 0.0;   // src/stomp_parameters.cpp:56; > 1e-9 turns the inverse-dynamics term on
  // PolicyImprovementLoop::readParameters (src/policy_improvement_loop.cpp:112-123)
  int num_rollouts = 10, num_reused_rollouts = 5, num_time_steps = 99;
  std::vector<double> noise_stddev, noise_decay;
  bool use_cumulative_costs = true;
  // engine-only knobs
  int dtype = STOMP_F64, device = 0;
  uint64_t seed = 0x57012011ull;
};

// ---- outputs of StompRobotModel / StompCollisionSpace that the path consumes ------------------------------
struct StompRobotModel {
  std::vector<stomp_segment> segments;  // KDL tree in DFS pre-order
  int reference_segment = 0;
  std::vector<stomp_sphere> collision_points;
  std::vector<stomp_joint_limit> joint_limits;  // one per planning-group joint
  // StompPlanningGroup::kdl_chain_ / id_solver_ (src/stomp_robot_model.cpp:181-185): one inertia per segment and the
  // root / tip of the inverse-dynamics chain; only read when StompParameters::torque_cost_weight > 1e-9
  std::vector<stomp_link_inertia> link_inertias;
  int chain_root_segment = -1, chain_tip_segment = -1;
  int getNumJoints() const { return int(joint_limits.size()); }
};

struct StompCollisionSpace {
  std::vector<uint8_t> voxels;  // squared cell distances, x-major
  int nx = 0, ny = 0, nz = 0, voxel_dtype = STOMP_VOXEL_U8_SQ;
  double origin[3] = {0, 0, 0}, resolution = 0.015;
  // what StompCollisionSpace::setStartState rebuilds the field from (src/stomp_collision_space.cpp:154-197): the environment's
  // collision objects and collision-map points, and the robot's own bodies outside the planning group at the start state
  // (StompRobotModelUrdf::bodiesAtState).  When `size` is set, StompOptimizer builds the field on the GPU
  // (stomp_engine_build_sdf_meshes) instead of uploading `voxels`.
  double size[3] = {0, 0, 0}, max_distance = 0.0;   // collision_space/size_*, max_propagation_distance
  std::vector<stomp_box> boxes;
  std::vector<stomp_cylinder> cylinders;
  std::vector<double> points;                       // [n][3]
  std::vector<stomp_body> bodies;
  std::vector<stomp_mesh_body> meshes;              // mesh collision objects / mesh link geometry (meshBodiesAtState)
  bool built_on_device() const { return size[0] > 0.0 && size[1] > 0.0 && size[2] > 0.0 && max_distance > 0.0; }
};

// StompTrajectory::fillInMinJerk (src/stomp_trajectory.cpp:179-223): quintic between the fixed points before / after the
// free block with zero start / end velocity and acceleration.  trajectory[d][t], t = 0 .. num_free_points-1.
inline void fillInMinJerk(const VectorXd& start, const VectorXd& goal, int num_free_points, double discretization,
                          std::vector<VectorXd>& trajectory) {
  double T[6];
  T[0] = 1.0;
  T[1] = (num_free_points + 1) * discretization;
  for (int i = 2; i <= 5; ++i) T[i] = T[i - 1] * T[1];
  trajectory.assign(start.size(), VectorXd(num_free_points, 0.0));
  for (size_t j = 0; j < start.size(); ++j) {
    const double x0 = start[j], x1 = goal[j];
    const double coeff[6] = {x0, 0, 0, (-20 * x0 + 20 * x1) / (2 * T[3]), (30 * x0 - 30 * x1) / (2 * T[4]), (-12 * x0 + 12 * x1) / (2 * T[5])};
    for (int i = 1; i <= num_free_points; ++i) {
      double t[6];
      t[0] = 1.0;
      t[1] = i * discretization;
      for (int k = 2; k <= 5; ++k) t[k] = t[k - 1] * t[1];
      double v = 0.0;
      for (int k = 0; k <= 5; ++k) v += t[k] * coeff[k];
      trajectory[j][i - 1] = v;
    }
  }
}

// ---- request pre / post-processing of StompPlannerNode::planKinematicPath (host side, either side of the path) ----------
// angles::shortest_angular_distance(from, to): (to - from) normalised into (-pi, pi]
inline double shortestAngularDistance(double from, double to) {
  double a = std::fmod(std::fmod(to - from, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI);   // normalize_angle_positive
  if (a > M_PI) a -= 2.0 * M_PI;
  return a;
}
// "fix the goal to move the shortest angular distance for wrap-around joints" (src/stomp_planner_node.cpp:207-217): a joint
// without limits is a continuous one (StompJoint::wrap_around_, src/stomp_robot_model.cpp:160-162)
inline void fixGoalForWrapAroundJoints(const VectorXd& start, VectorXd& goal, const std::vector<stomp_joint_limit>& limits) {
  for (size_t j = 0; j < goal.size() && j < limits.size(); ++j)
    if (!limits[j].has_limits) goal[j] = start[j] + shortestAngularDistance(start[j], goal[j]);
}
// time_from_start of the response trajectory (src/stomp_planner_node.cpp:257-279): points = start, the N optimised points,
// goal; every step lasts the trajectory discretization unless a joint would exceed its velocity limit
// (joint_velocity_limits/<joint>, default unlimited), then as long as that joint needs.  trajectory[d][t], t = 0 .. N-1.
inline std::vector<double> timeFromStart(const VectorXd& start, const std::vector<VectorXd>& trajectory, const VectorXd& goal,
                                         double discretization, const std::vector<double>& velocity_limits) {
  const size_t D = trajectory.size(), N = D ? trajectory[0].size() : 0;
  auto position = [&](size_t i, size_t j) { return i == 0 ? start[j] : (i == N + 1 ? goal[j] : trajectory[j][i - 1]); };
  std::vector<double> t(N + 2, 0.0);
  for (size_t i = 1; i < N + 2; ++i) {
    double duration = discretization;
    for (size_t j = 0; j < D; ++j) {
      const double limit = j < velocity_limits.size() ? velocity_limits[j] : std::numeric_limits<double>::max();
      const double d = std::fabs(position(i, j) - position(i - 1, j)) / limit;
      if (d > duration) duration = d;
    }
    t[i] = t[i - 1] + duration;
  }
  return t;
}

// shared RAII owner of one engine handle
class Engine {
 public:
  explicit Engine(void* h) : h_(h) {}
  ~Engine() { if (h_) stomp_engine_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  void* get() const { return h_; }
  int D = 0, N = 0, R = 0, R_reused = 0;
 private:
  void* h_;
};

// ---- Policy / Task: the plugin interfaces ------------------------------------------------------------------
class Policy {
 public:
  virtual ~Policy() {}
  virtual bool setNumTimeSteps(const int num_time_steps) = 0;
  virtual bool getNumTimeSteps(int& num_time_steps) = 0;
  virtual bool getNumDimensions(int& num_dimensions) = 0;
  virtual bool getNumParameters(std::vector<int>& num_params) = 0;
  virtual bool getBasisFunctions(std::vector<MatrixXd>& basis_functions) = 0;
  virtual bool getControlCosts(std::vector<MatrixXd>& control_costs) = 0;
  virtual bool updateParameters(const std::vector<MatrixXd>& updates) = 0;
  virtual bool getParameters(std::vector<VectorXd>& parameters) = 0;
  virtual bool setParameters(const std::vector<VectorXd>& parameters) = 0;
  virtual bool computeControlCosts(const std::vector<MatrixXd>& control_cost_matrices, const std::vector<VectorXd>& parameters,
                                   const std::vector<VectorXd>& noise, const double weight,
                                   std::vector<VectorXd>& control_costs) = 0;
};

class Task {
 public:
  virtual ~Task() {}
  virtual bool initialize(const StompParameters& params, int num_time_steps) = 0;
  // executes one rollout: parameters[d][t] -> costs[t]   (include/stomp_motion_planner/task.h:70)
  virtual bool execute(std::vector<VectorXd>& parameters, VectorXd& costs, const int iteration_number) = 0;
  virtual bool getPolicy(std::shared_ptr<Policy>& policy) = 0;
  virtual bool setPolicy(const std::shared_ptr<Policy> policy) = 0;
  virtual bool getControlCostWeight(double& control_cost_weight) = 0;
  // ---- optional extensions (defaults keep an existing cost plugin source compatible: the five methods above are the
  //      reference's whole Task interface, include/stomp_motion_planner/task.h:52-93) -------------------------------------
  // All rollouts of an iteration at once: rollouts[r][d][t] -> costs(r, t).  The default calls execute() rollout by rollout
  // like PolicyImprovementLoop::runSingleIteration does (src/policy_improvement_loop.cpp:160-166); a plugin that evaluates on
  // the GPU overrides it and saves R - 1 host round trips per iteration.
  virtual bool executeBatch(std::vector<std::vector<VectorXd> >& rollouts, MatrixXd& costs, const int iteration_number) {
    VectorXd tmp;
    for (size_t r = 0; r < rollouts.size(); ++r) {
      if (!execute(rollouts[r], tmp, iteration_number)) return false;
      if (costs.rows() != int(rollouts.size()) || costs.cols() != int(tmp.size())) costs = MatrixXd(int(rollouts.size()), int(tmp.size()));
      for (size_t t = 0; t < tmp.size(); ++t) costs(int(r), int(t)) = tmp[t];
    }
    return true;
  }
  // Engine the PI^2 state lives on.  The loop asks the task's policy (CovariantTrajectoryPolicy::getEngine) when this returns
  // null, so a custom Task does not need to know about engines at all.
  virtual std::shared_ptr<Engine> getEngine() { return nullptr; }
};

// ---- CovariantTrajectoryPolicy ---------------------------------------------------------------------------------
class CovariantTrajectoryPolicy : public Policy {
 public:
  explicit CovariantTrajectoryPolicy(std::shared_ptr<Engine> engine) : e_(engine) {}
  std::shared_ptr<Engine> getEngine() const { return e_; }
  // setToMinControlCost (src/covariant_trajectory_policy.cpp:102-112): also resets the rollout-reuse state
  bool setToMinControlCost(const VectorXd& start, const VectorXd& goal) {
    return stomp_engine_set_problems(e_->get(), start.data(), goal.data()) == 0;
  }
  bool setNumTimeSteps(const int n) override { return n == e_->N; }
  bool getNumTimeSteps(int& n) override { n = e_->N; return true; }
  bool getNumDimensions(int& d) override { d = e_->D; return true; }
  bool getNumParameters(std::vector<int>& n) override { n.assign(e_->D, e_->N); return true; }
  bool getBasisFunctions(std::vector<MatrixXd>& basis) override {  // identity (params == time steps)
    MatrixXd I(e_->N, e_->N);
    for (int i = 0; i < e_->N; ++i) I(i, i) = 1.0;
    basis.assign(e_->D, I);
    return true;
  }
  bool getControlCosts(std::vector<MatrixXd>& control_costs) override { return getMatrix(STOMP_FIELD_CONTROL_COST, control_costs); }
  bool getInvControlCosts(std::vector<MatrixXd>& inv) { return getMatrix(STOMP_FIELD_INV_CONTROL_COST, inv); }
  bool updateParameters(const std::vector<MatrixXd>& updates) override {  // row 0 only, like the reference
    std::vector<double> u(size_t(e_->D) * e_->N);
    if (int(updates.size()) != e_->D) return false;
    for (int d = 0; d < e_->D; ++d)
      for (int t = 0; t < e_->N; ++t) u[size_t(d) * e_->N + t] = updates[d](0, t);
    return stomp_engine_update_parameters(e_->get(), u.data()) == 0;
  }
  bool getParameters(std::vector<VectorXd>& parameters) override {
    std::vector<double> th(size_t(e_->D) * e_->N);
    if (stomp_engine_get_parameters(e_->get(), th.data())) return false;
    parameters.resize(e_->D);
    for (int d = 0; d < e_->D; ++d) parameters[d].assign(th.begin() + size_t(d) * e_->N, th.begin() + size_t(d + 1) * e_->N);
    return true;
  }
  bool setParameters(const std::vector<VectorXd>& parameters) override {
    std::vector<double> th;
    for (const VectorXd& p : parameters) th.insert(th.end(), p.begin(), p.end());
    return th.size() == size_t(e_->D) * e_->N && stomp_engine_set_parameters(e_->get(), th.data()) == 0;
  }
  bool computeControlCosts(const std::vector<MatrixXd>&, const std::vector<VectorXd>& parameters, const std::vector<VectorXd>& noise,
                           const double weight, std::vector<VectorXd>& control_costs) override {
    std::vector<double> p, n, c(size_t(e_->D) * e_->N);
    for (int d = 0; d < e_->D; ++d) p.insert(p.end(), parameters[d].begin(), parameters[d].end()), n.insert(n.end(), noise[d].begin(), noise[d].end());
    if (stomp_engine_compute_control_costs(e_->get(), p.data(), n.data(), 1, weight, c.data())) return false;
    control_costs.resize(e_->D);
    for (int d = 0; d < e_->D; ++d) control_costs[d].assign(c.begin() + size_t(d) * e_->N, c.begin() + size_t(d + 1) * e_->N);
    return true;
  }
  std::shared_ptr<Engine> engine() const { return e_; }

 private:
  bool getMatrix(int field, std::vector<MatrixXd>& out) {
    MatrixXd m(e_->N, e_->N);
    if (stomp_engine_get(e_->get(), field, m.data.data(), m.data.size() * sizeof(double))) return false;
    out.assign(e_->D, m);
    return true;
  }
  std::shared_ptr<Engine> e_;
};
typedef CovariantTrajectoryPolicy CovariantMovementPrimitive;

// ---- PolicyImprovement ---------------------------------------------------------------------------------------
class PolicyImprovement {
 public:
  bool initialize(const int num_rollouts, const int num_time_steps, const int num_reused_rollouts, const int num_extra_rollouts,
                  std::shared_ptr<Policy> policy, bool use_cumulative_costs = true) {
    (void)use_cumulative_costs;
    auto ctp = std::dynamic_pointer_cast<CovariantTrajectoryPolicy>(policy);
    if (!ctp) return false;  // the PI^2 state lives on the policy's engine
    e_ = ctp->engine();
    if (num_extra_rollouts != 1) return false;  // hard-coded 1 in the reference loop (policy_improvement_loop.cpp:103)
    return (initialized_ = (num_rollouts == e_->R && num_time_steps == e_->N && num_reused_rollouts == e_->R_reused));
  }
  bool getRollouts(std::vector<std::vector<VectorXd> >& rollouts, const std::vector<double>& noise_stddev) {
    if (!initialized_) return false;
    std::vector<double> buf(size_t(e_->R) * e_->D * e_->N);
    int32_t gen = 0;
    if (stomp_engine_get_rollouts(e_->get(), noise_stddev.data(), buf.data(), &gen)) return false;
    rollouts.assign(gen, std::vector<VectorXd>(e_->D));
    for (int r = 0; r < gen; ++r)
      for (int d = 0; d < e_->D; ++d) {
        const double* p = &buf[(size_t(r) * e_->D + d) * e_->N];
        rollouts[r][d].assign(p, p + e_->N);
      }
    num_gen_ = gen;
    return true;
  }
  // costs: [num generated rollouts][num_time_steps]
  bool setRolloutCosts(const MatrixXd& costs, const double control_cost_weight, std::vector<double>& rollout_costs_total) {
    if (!initialized_ || costs.rows() < num_gen_ || costs.cols() != e_->N) return false;
    rollout_costs_total.resize(e_->R);
    return stomp_engine_set_rollout_costs(e_->get(), costs.data.data(), control_cost_weight, rollout_costs_total.data()) == 0;
  }
  bool improvePolicy(std::vector<MatrixXd>& parameter_updates) {
    if (!initialized_) return false;
    std::vector<double> u(size_t(e_->D) * e_->N);
    if (stomp_engine_improve_policy(e_->get(), u.data())) return false;
    parameter_updates.assign(e_->D, MatrixXd(e_->N, e_->N));  // only row 0 is ever filled (policy_improvement.cpp:370-383)
    for (int d = 0; d < e_->D; ++d)
      for (int t = 0; t < e_->N; ++t) parameter_updates[d](0, t) = u[size_t(d) * e_->N + t];
    return true;
  }
  // the single extra rollout must be the policy's current parameters (as in policy_improvement_loop.cpp:186-192)
  bool addExtraRollouts(std::vector<std::vector<VectorXd> >& rollouts, std::vector<VectorXd>& rollout_costs) {
    if (!initialized_ || rollouts.size() != 1 || rollout_costs.size() != 1) return false;
    return stomp_engine_add_extra_rollouts(e_->get(), rollout_costs[0].data()) == 0;
  }

 private:
  std::shared_ptr<Engine> e_;
  bool initialized_ = false;
  int num_gen_ = 0;
};

class StompOptimizer;

// ---- PolicyImprovementLoop -------------------------------------------------------------------------------------
class PolicyImprovementLoop {
 public:
  bool initialize(const StompParameters& params, std::shared_ptr<Task> task);
  bool runSingleIteration(const int iteration_number);
  // results of the noise-less rollout of the last iteration
  double lastNoiselessCost() const { return last_cost_; }
  bool lastNoiselessCollisionFree() const { return last_collision_free_; }

 private:
  StompParameters params_;
  std::shared_ptr<Task> task_;
  std::shared_ptr<Policy> policy_;
  std::shared_ptr<Engine> e_;
  PolicyImprovement policy_improvement_;
  double control_cost_weight_ = 0.0, last_cost_ = 0.0;
  bool last_collision_free_ = false, fused_ = false, initialized_ = false;
};

// ---- StompOptimizer: the built-in GPU Task + the outer optimisation loop ------------------------------------------
struct STOMPStatistics {  // msg/STOMPStatistics.msg
  bool success = false;
  int success_iteration = -1, collision_success_iteration = -1;
  double best_cost = 0.0;
  std::vector<double> costs;
};

class StompOptimizer : public Task, public std::enable_shared_from_this<StompOptimizer> {
 public:
  // start / goal: the fixed trajectory end points of the planning group (free_vars_start_-1, free_vars_end_+1)
  StompOptimizer(const VectorXd& start, const VectorXd& goal, const StompRobotModel* robot_model,
                 const StompParameters* parameters, const StompCollisionSpace* collision_space)
      : start_(start), goal_(goal), robot_model_(robot_model), parameters_(parameters), collision_space_(collision_space) {
    ok_ = initializeEngine();
  }
  bool ok() const { return ok_; }

  bool initialize(const StompParameters&, int num_time_steps) override { return ok_ && num_time_steps == engine_->N; }
  bool execute(std::vector<VectorXd>& parameters, VectorXd& costs, const int iteration_number) override {
    std::vector<double> p;
    for (const VectorXd& v : parameters) p.insert(p.end(), v.begin(), v.end());
    if (p.size() != size_t(engine_->D) * engine_->N) return false;
    costs.assign(engine_->N, 0.0);
    int32_t cf = 0;
    if (stomp_engine_execute(engine_->get(), p.data(), 1, iteration_number, costs.data(), &cf)) return false;
    last_trajectory_collision_free_ = cf != 0;
    last_trajectory_cost_ = 0.0;
    for (double c : costs) last_trajectory_cost_ += c;
    return true;
  }
  // all rollouts in one stomp_engine_execute call (one upload, one k_cost launch, one read-back)
  bool executeBatch(std::vector<std::vector<VectorXd> >& rollouts, MatrixXd& costs, const int iteration_number) override {
    const size_t DN = size_t(engine_->D) * engine_->N;
    std::vector<double> p;
    p.reserve(rollouts.size() * DN);
    for (const std::vector<VectorXd>& ro : rollouts)
      for (const VectorXd& v : ro) p.insert(p.end(), v.begin(), v.end());
    if (rollouts.empty() || p.size() != rollouts.size() * DN) return false;
    costs = MatrixXd(int(rollouts.size()), engine_->N);
    std::vector<int32_t> cf(rollouts.size(), 0);
    if (stomp_engine_execute(engine_->get(), p.data(), int32_t(rollouts.size()), iteration_number, costs.data.data(), cf.data())) return false;
    last_trajectory_collision_free_ = cf.back() != 0;
    return true;
  }
  bool getPolicy(std::shared_ptr<Policy>& policy) override { policy = policy_; return true; }
  bool setPolicy(const std::shared_ptr<Policy>) override { return true; }
  bool getControlCostWeight(double& w) override { w = parameters_->smoothness_cost_weight; return true; }
  std::shared_ptr<Engine> getEngine() override { return engine_; }

  // StompOptimizer::optimize (src/stomp_optimizer.cpp:249-401), STOMP branch: the iteration loop and its bookkeeping
  // (collision_free_iteration_, success iterations, best trajectory, early exit) run on the device.
  bool optimize(STOMPStatistics* stats = nullptr) {
    if (!ok_) return false;
    const int max_it = parameters_->max_iterations;
    STOMPStatistics st;
    st.costs.assign(max_it, 0.0);
    int32_t success = 0, success_it = -1, coll_it = -1, last_imp = -1, iterations = 0;
    double best_cost = 0.0;
    stomp_optimize_stats os = {&success, &success_it, &coll_it, &last_imp, &iterations, &best_cost, st.costs.data()};
    if (stomp_engine_optimize(engine_->get(), max_it, parameters_->max_iterations_after_collision_free, &os)) return false;
    st.costs.resize(iterations);
    st.success = success != 0;
    st.success_iteration = success_it;
    st.collision_success_iteration = coll_it;
    st.best_cost = best_cost;
    last_improvement_iteration_ = last_imp;
    iteration_ = iterations;
    if (!st.costs.empty()) last_trajectory_cost_ = st.costs.back();
    if (stats) *stats = st;
    return true;
  }
  // group_trajectory_ after optimize(): the best trajectory (after joint-limit handling), [D][N]
  bool getBestTrajectory(std::vector<VectorXd>& trajectory) {
    std::vector<double> buf(size_t(engine_->D) * engine_->N);
    if (stomp_engine_get(engine_->get(), STOMP_FIELD_BEST_TRAJECTORY, buf.data(), buf.size() * sizeof(double))) return false;
    trajectory.resize(engine_->D);
    for (int d = 0; d < engine_->D; ++d) trajectory[d].assign(buf.begin() + size_t(d) * engine_->N, buf.begin() + size_t(d + 1) * engine_->N);
    return true;
  }

  double getLastTrajectoryCost() const { return last_trajectory_cost_; }
  bool isLastTrajectoryCollisionFree() const { return last_trajectory_collision_free_; }
  int getLastImprovementIteration() const { return last_improvement_iteration_; }

 private:
  bool initializeEngine() {
    stomp_engine_desc d = {};
    d.num_dimensions = robot_model_->getNumJoints();
    d.num_time_steps = parameters_->num_time_steps;
    d.num_rollouts = parameters_->num_rollouts;
    d.num_reused_rollouts = parameters_->num_reused_rollouts;
    d.num_problems = 1;
    d.dtype = parameters_->dtype;
    d.use_cumulative_costs = parameters_->use_cumulative_costs ? 1 : 0;
    d.sdf_mode = STOMP_SDF_NEAREST;
    d.device = parameters_->device;
    d.rollout_shard_rank = 0;
    d.rollout_shard_world = 1;
    // group trajectory: N free + 2*6 padded points; StompTrajectory::getDuration() returns int and seeds the
    // policy's movement duration (stomp_trajectory.h:141, stomp_optimizer.cpp:187): keep the truncation
    d.movement_duration = double(int((parameters_->num_time_steps + 2 * (STOMP_DIFF_RULE_LENGTH - 1) - 1) *
                                     parameters_->trajectory_discretization));
    d.discretization = parameters_->trajectory_discretization;
    d.derivative_costs[0] = parameters_->smoothness_cost_velocity;
    d.derivative_costs[1] = parameters_->smoothness_cost_acceleration;
    d.derivative_costs[2] = parameters_->smoothness_cost_jerk;
    d.ridge_factor = parameters_->ridge_factor;
    d.smoothness_cost_weight = parameters_->smoothness_cost_weight;
    d.obstacle_cost_weight = parameters_->obstacle_cost_weight;
    void* h = nullptr;
    if (stomp_engine_create(&d, &h)) return false;
    engine_ = std::make_shared<Engine>(h);
    engine_->D = d.num_dimensions; engine_->N = d.num_time_steps; engine_->R = d.num_rollouts; engine_->R_reused = d.num_reused_rollouts;
    if (stomp_engine_set_robot(h, robot_model_->segments.data(), int(robot_model_->segments.size()), robot_model_->reference_segment,
                               robot_model_->collision_points.data(), int(robot_model_->collision_points.size()),
                               robot_model_->joint_limits.data()))
      return false;
    if (collision_space_->built_on_device()) {
      const StompCollisionSpace& cs = *collision_space_;
      if (stomp_engine_build_sdf_meshes(h, cs.size, cs.origin, cs.resolution, cs.max_distance, cs.boxes.data(), int(cs.boxes.size()),
                                        cs.cylinders.data(), int(cs.cylinders.size()), cs.points.data(), int64_t(cs.points.size() / 3),
                                        cs.bodies.data(), int(cs.bodies.size()), cs.meshes.data(), int(cs.meshes.size())))
        return false;
    } else if (stomp_engine_set_sdf(h, collision_space_->voxels.data(), collision_space_->nx, collision_space_->ny, collision_space_->nz,
                                    collision_space_->origin, collision_space_->resolution, collision_space_->voxel_dtype)) {
      return false;
    }
    if (parameters_->torque_cost_weight > 1e-9) {   // src/stomp_optimizer.cpp:1120
      const double gravity[3] = {0.0, 0.0, -9.8};   // src/stomp_robot_model.cpp:184
      if (robot_model_->link_inertias.size() != robot_model_->segments.size() ||
          stomp_engine_set_dynamics(h, robot_model_->link_inertias.data(), robot_model_->chain_root_segment,
                                    robot_model_->chain_tip_segment, gravity, parameters_->torque_cost_weight))
        return false;
    }
    std::vector<double> sd = parameters_->noise_stddev, dc = parameters_->noise_decay;
    sd.resize(d.num_dimensions, 2.0);
    dc.resize(d.num_dimensions, 0.999);
    if (stomp_engine_set_noise(h, sd.data(), dc.data()) || stomp_engine_seed(h, parameters_->seed)) return false;
    policy_ = std::make_shared<CovariantTrajectoryPolicy>(engine_);
    return policy_->setToMinControlCost(start_, goal_);
  }

  VectorXd start_, goal_;
  const StompRobotModel* robot_model_;
  const StompParameters* parameters_;
  const StompCollisionSpace* collision_space_;
  std::shared_ptr<Engine> engine_;
  std::shared_ptr<CovariantTrajectoryPolicy> policy_;
  bool ok_ = false, last_trajectory_collision_free_ = false;
  double last_trajectory_cost_ = 0.0;
  int iteration_ = 0, last_improvement_iteration_ = -1;
};

// ---- PolicyImprovementLoop implementation -----------------------------------------------------------------------
inline bool PolicyImprovementLoop::initialize(const StompParameters& params, std::shared_ptr<Task> task) {
  params_ = params;
  task_ = task;
  if (!task_->initialize(params_, params_.num_time_steps)) return false;
  if (!task_->getPolicy(policy_) || !task_->getControlCostWeight(control_cost_weight_)) return false;
  e_ = task_->getEngine();
  if (!e_)
    if (CovariantTrajectoryPolicy* ctp = dynamic_cast<CovariantTrajectoryPolicy*>(policy_.get())) e_ = ctp->getEngine();
  if (!e_) return false;
  params_.noise_stddev.resize(e_->D, 2.0);
  params_.noise_decay.resize(e_->D, 0.999);
  if (!policy_improvement_.initialize(params_.num_rollouts, params_.num_time_steps, params_.num_reused_rollouts, 1, policy_,
                                      params_.use_cumulative_costs))
    return false;
  fused_ = dynamic_cast<StompOptimizer*>(task_.get()) != nullptr;
  return (initialized_ = true);
}

inline bool PolicyImprovementLoop::runSingleIteration(const int iteration_number) {
  if (!initialized_) return false;
  if (fused_) {  // the built-in GPU task: the whole iteration stays on the device
    double cost = 0.0;
    int32_t cf = 0;
    stomp_iter_stats st = {&cost, &cf, 0, 0, nullptr};
    if (stomp_engine_iterate(e_->get(), iteration_number, &st)) return false;
    last_cost_ = cost;
    last_collision_free_ = cf != 0;
    return true;
  }
  // any other Task: the reference's call sequence (src/policy_improvement_loop.cpp:143-202)
  std::vector<double> noise(e_->D);
  for (int i = 0; i < e_->D; ++i) noise[i] = params_.noise_stddev[i] * std::pow(params_.noise_decay[i], iteration_number - 1);
  std::vector<std::vector<VectorXd> > rollouts;
  if (!policy_improvement_.getRollouts(rollouts, noise)) return false;
  MatrixXd rollout_costs(int(rollouts.size()), e_->N);
  VectorXd tmp;
  if (!task_->executeBatch(rollouts, rollout_costs, iteration_number)) return false;
  if (rollout_costs.rows() != int(rollouts.size()) || rollout_costs.cols() != e_->N) return false;
  std::vector<double> all_costs;
  if (!policy_improvement_.setRolloutCosts(rollout_costs, control_cost_weight_, all_costs)) return false;
  std::vector<MatrixXd> parameter_updates;
  if (!policy_improvement_.improvePolicy(parameter_updates)) return false;
  if (!policy_->updateParameters(parameter_updates)) return false;
  std::vector<VectorXd> parameters;
  if (!policy_->getParameters(parameters)) return false;
  if (!task_->execute(parameters, tmp, iteration_number)) return false;
  last_cost_ = 0.0;
  for (double c : tmp) last_cost_ += c;
  std::vector<std::vector<VectorXd> > extra_rollout(1, parameters);
  std::vector<VectorXd> extra_rollout_cost(1, tmp);
  return policy_improvement_.addExtraRollouts(extra_rollout, extra_rollout_cost);
}

}  // namespace stomp_motion_planner
