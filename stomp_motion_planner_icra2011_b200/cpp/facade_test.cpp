// facade_test.cpp — exercises the reference-shaped C++ classes above the C ABI on a small synthetic arm:
//   1. StompOptimizer::optimize() (fused GPU iteration) lowers the noise-less trajectory cost;
//   2. PolicyImprovementLoop with a user-defined Task (the reference's call sequence, cost plugin called once
//      per rollout from the host) produces the same policy as the fused path for the same seed.
// Needs a CUDA device (run by tests/test_gpu_facade.py).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "stomp_motion_planner/stomp_b200_facade.hpp"

using namespace stomp_motion_planner;

static stomp_segment seg(int parent, int type, int group, double px, double py, double pz, double ax, double ay, double az) {
  stomp_segment s;
  std::memset(&s, 0, sizeof(s));
  s.parent = parent; s.joint_type = type; s.group_index = group;
  s.rot[0] = s.rot[4] = s.rot[8] = 1.0;
  s.pos[0] = px; s.pos[1] = py; s.pos[2] = pz;
  s.axis[0] = ax; s.axis[1] = ay; s.axis[2] = az;
  return s;
}

// forwards execute() to an inner StompOptimizer: forces PolicyImprovementLoop onto the generic Task path
class ForwardingTask : public Task {
 public:
  explicit ForwardingTask(std::shared_ptr<StompOptimizer> inner) : inner_(inner) {}
  bool initialize(const StompParameters& p, int n) override { return inner_->initialize(p, n); }
  bool execute(std::vector<VectorXd>& parameters, VectorXd& costs, const int it) override {
    ++calls;
    return inner_->execute(parameters, costs, it);
  }
  bool getPolicy(std::shared_ptr<Policy>& policy) override { return inner_->getPolicy(policy); }
  bool setPolicy(const std::shared_ptr<Policy> policy) override { return inner_->setPolicy(policy); }
  bool getControlCostWeight(double& w) override { return inner_->getControlCostWeight(w); }
  // no getEngine(): only the reference's five Task methods, like an existing cost plugin; the loop finds the engine through
  // the policy
  int calls = 0;
 protected:
  std::shared_ptr<StompOptimizer> inner_;
};

// the optional batched hook: all rollouts of an iteration in one call
class BatchedTask : public ForwardingTask {
 public:
  using ForwardingTask::ForwardingTask;
  bool executeBatch(std::vector<std::vector<VectorXd> >& rollouts, MatrixXd& costs, const int it) override {
    ++batch_calls;
    return inner_->executeBatch(rollouts, costs, it);
  }
  int batch_calls = 0;
};

int main() {
  // 4-DOF arm: z, y, y, x axes
  StompRobotModel robot;
  robot.segments.push_back(seg(-1, STOMP_JOINT_FIXED, -1, 0, 0, 0, 0, 0, 1));
  robot.segments.push_back(seg(0, STOMP_JOINT_REVOLUTE, 0, 0, 0, 0.5, 0, 0, 1));
  robot.segments.push_back(seg(1, STOMP_JOINT_REVOLUTE, 1, 0.05, 0, 0.1, 0, 1, 0));
  robot.segments.push_back(seg(2, STOMP_JOINT_REVOLUTE, 2, 0.35, 0, 0, 0, 1, 0));
  robot.segments.push_back(seg(3, STOMP_JOINT_REVOLUTE, 3, 0.30, 0, 0, 1, 0, 0));
  robot.segments.push_back(seg(4, STOMP_JOINT_FIXED, -1, 0.15, 0, 0, 0, 0, 1));
  for (int link = 2; link <= 4; ++link)
    for (int i = 0; i < 5; ++i) {
      stomp_sphere sp;
      std::memset(&sp, 0, sizeof(sp));
      sp.segment = link; sp.radius = 0.05; sp.clearance = 0.07;
      sp.pos[0] = robot.segments[link + 1].pos[0] * i / 4.0;
      robot.collision_points.push_back(sp);
    }
  robot.joint_limits = {{1, 0, -2.5, 2.5}, {1, 0, -1.5, 1.5}, {1, 0, -2.2, 2.2}, {0, 0, 0, 0}};

  // distance field: a box obstacle in front of the arm, brute-force squared cell distances capped at 12^2
  StompCollisionSpace space;
  space.nx = 60; space.ny = 60; space.nz = 60; space.resolution = 0.03;
  space.origin[0] = -0.9; space.origin[1] = -0.9; space.origin[2] = -0.2;
  space.voxels.assign(size_t(60) * 60 * 60, 144);
  const int cap = 12;
  auto occupied = [&](int x, int y, int z) {
    double wx = space.origin[0] + x * space.resolution, wy = space.origin[1] + y * space.resolution, wz = space.origin[2] + z * space.resolution;
    return std::fabs(wx - 0.55) < 0.08 && std::fabs(wy - 0.0) < 0.25 && std::fabs(wz - 0.55) < 0.12;
  };
  for (int x = 0; x < 60; ++x) for (int y = 0; y < 60; ++y) for (int z = 0; z < 60; ++z) {
    if (!occupied(x, y, z)) continue;
    for (int dx = -cap; dx <= cap; ++dx) for (int dy = -cap; dy <= cap; ++dy) for (int dz = -cap; dz <= cap; ++dz) {
      int X = x + dx, Y = y + dy, Z = z + dz;
      if (X < 0 || Y < 0 || Z < 0 || X >= 60 || Y >= 60 || Z >= 60) continue;
      int d2 = dx * dx + dy * dy + dz * dz;
      uint8_t& v = space.voxels[(size_t(X) * 60 + Y) * 60 + Z];
      if (d2 < v) v = uint8_t(d2);
    }
  }

  StompParameters params;
  params.num_time_steps = 60; params.num_rollouts = 10; params.num_reused_rollouts = 5;
  params.max_iterations = 40; params.max_iterations_after_collision_free = 1000;
  params.smoothness_cost_weight = 1e-6; params.use_cumulative_costs = false;
  params.noise_stddev.assign(4, 2.0); params.noise_decay.assign(4, 0.999);
  VectorXd start = {-0.6, 0.3, 0.4, 0.0}, goal = {0.6, 0.3, 0.4, 0.5};

  // ---- 1. optimize() -----------------------------------------------------------------------------------------
  auto opt = std::make_shared<StompOptimizer>(start, goal, &robot, &params, &space);
  if (!opt->ok()) { std::printf("FAIL create: %s\n", lastError()); return 1; }
  STOMPStatistics stats;
  if (!opt->optimize(&stats)) { std::printf("FAIL optimize: %s\n", lastError()); return 1; }
  std::printf("optimize: first cost %.6f, best cost %.6f, %zu iterations, collision-free at %d\n", stats.costs.front(),
              stats.best_cost, stats.costs.size(), stats.collision_success_iteration);
  if (!(stats.best_cost <= stats.costs.front()) || stats.costs.size() != 40u) { std::printf("FAIL cost did not improve\n"); return 1; }
  std::vector<VectorXd> best;
  if (!opt->getBestTrajectory(best) || best.size() != 4u || best[0].size() != 60u) { std::printf("FAIL best trajectory\n"); return 1; }
  for (int d = 0; d < 3; ++d)
    for (double v : best[d])
      if (v > robot.joint_limits[d].max + 1e-4 || v < robot.joint_limits[d].min - 1e-4) { std::printf("FAIL joint limits\n"); return 1; }

  // ---- 2. fused path == reference call sequence with a host Task --------------------------------------------------
  auto a = std::make_shared<StompOptimizer>(start, goal, &robot, &params, &space);
  auto b_inner = std::make_shared<StompOptimizer>(start, goal, &robot, &params, &space);
  auto b = std::make_shared<ForwardingTask>(b_inner);
  auto c_inner = std::make_shared<StompOptimizer>(start, goal, &robot, &params, &space);
  auto c = std::make_shared<BatchedTask>(c_inner);
  PolicyImprovementLoop loop_a, loop_b, loop_c;
  if (!loop_a.initialize(params, a) || !loop_b.initialize(params, b) || !loop_c.initialize(params, c)) { std::printf("FAIL loop init: %s\n", lastError()); return 1; }
  double max_diff = 0.0;
  for (int it = 1; it <= 6; ++it) {
    if (!loop_a.runSingleIteration(it) || !loop_b.runSingleIteration(it) || !loop_c.runSingleIteration(it)) { std::printf("FAIL iteration: %s\n", lastError()); return 1; }
    std::shared_ptr<Policy> pa, pb, pc;
    a->getPolicy(pa); b->getPolicy(pb); c->getPolicy(pc);
    std::vector<VectorXd> ta, tb, tc;
    pa->getParameters(ta); pb->getParameters(tb); pc->getParameters(tc);
    for (size_t d = 0; d < ta.size(); ++d)
      for (size_t t = 0; t < ta[d].size(); ++t) {
        max_diff = std::fmax(max_diff, std::fabs(ta[d][t] - tb[d][t]));
        if (tb[d][t] != tc[d][t]) { std::printf("FAIL batched plugin hook differs from the per-rollout one\n"); return 1; }
      }
    if (std::fabs(loop_a.lastNoiselessCost() - loop_b.lastNoiselessCost()) > 1e-9 * (1.0 + std::fabs(loop_a.lastNoiselessCost()))) {
      std::printf("FAIL noise-less cost differs at iteration %d: %.12g vs %.12g\n", it, loop_a.lastNoiselessCost(), loop_b.lastNoiselessCost());
      return 1;
    }
  }
  // 10 rollouts at iteration 1, 5 afterwards, +1 noise-less each: 6*1 + 10 + 5*5 = 41 plugin calls
  std::printf("fused vs host-Task path: max |theta diff| = %.3e, host Task::execute calls = %d\n", max_diff, b->calls);
  if (max_diff > 1e-9 || b->calls != 41) { std::printf("FAIL paths disagree\n"); return 1; }
  // the batched hook: one executeBatch per iteration + one execute for the noise-less rollout
  if (c->batch_calls != 6 || c->calls != 6) { std::printf("FAIL batched hook call counts %d %d\n", c->batch_calls, c->calls); return 1; }

  // ---- 3. Policy interface bits ---------------------------------------------------------------------------------
  std::shared_ptr<Policy> pol;
  a->getPolicy(pol);
  std::vector<MatrixXd> cc;
  std::vector<int> np;
  int D = 0, N = 0;
  if (!pol->getControlCosts(cc) || !pol->getNumDimensions(D) || !pol->getNumTimeSteps(N) || !pol->getNumParameters(np)) return 1;
  if (D != 4 || N != 60 || cc.size() != 4u || cc[0].rows() != 60 || np[0] != 60 || !(cc[0](30, 30) > 0.0)) { std::printf("FAIL policy getters\n"); return 1; }
  CovariantMovementPrimitive* alias = dynamic_cast<CovariantMovementPrimitive*>(pol.get());
  if (!alias) { std::printf("FAIL alias\n"); return 1; }
  std::vector<VectorXd> mj;
  fillInMinJerk(start, goal, 60, 0.05, mj);
  const double mid = 0.5 * (start[0] + goal[0]);
  if (mj.size() != 4u || mj[0].size() != 60u || std::fabs(0.5 * (mj[0][29] + mj[0][30]) - mid) > 1e-9 ||
      std::fabs(mj[0][0] - start[0]) > 1e-3 || std::fabs(mj[0][59] - goal[0]) > 1e-3) { std::printf("FAIL min jerk\n"); return 1; }
  // ---- 4. setStartState: the field rebuilt on the device from a collision object + a robot body outside the group ----------
  {
    StompCollisionSpace built;
    built.size[0] = built.size[1] = built.size[2] = 1.8; built.max_distance = 0.36; built.resolution = 0.03;
    built.origin[0] = -0.9; built.origin[1] = -0.9; built.origin[2] = -0.2;
    stomp_box bx;
    std::memset(&bx, 0, sizeof(bx));
    bx.position[0] = 0.55; bx.position[2] = 0.55; bx.orientation[3] = 1.0;
    bx.dimensions[0] = 0.16; bx.dimensions[1] = 0.5; bx.dimensions[2] = 0.24;
    built.boxes.push_back(bx);
    stomp_body torso;                      // a "torso" column behind the arm: never touched, but present in the field
    std::memset(&torso, 0, sizeof(torso));
    torso.type = STOMP_BODY_CYLINDER; torso.dimensions[0] = 0.1; torso.dimensions[1] = 0.6;
    torso.position[0] = -0.5; torso.position[2] = 0.3; torso.orientation[3] = 1.0; torso.scale = 1.0; torso.padding = 0.01;
    built.bodies.push_back(torso);
    // a mesh collision object (bodies::ConvexMesh = the hull of its vertices): a cube given by its corners and one interior vertex
    std::vector<double> cube;
    for (int sx = -1; sx <= 1; sx += 2)
      for (int sy = -1; sy <= 1; sy += 2)
        for (int sz = -1; sz <= 1; sz += 2) { cube.push_back(0.1 * sx); cube.push_back(0.1 * sy); cube.push_back(0.1 * sz); }
    cube.push_back(0.01); cube.push_back(0.02); cube.push_back(-0.03);
    stomp_mesh_body mesh;
    std::memset(&mesh, 0, sizeof(mesh));
    mesh.vertices = cube.data(); mesh.num_vertices = int32_t(cube.size() / 3);
    mesh.position[0] = -0.452; mesh.position[1] = 0.601; mesh.position[2] = 1.003; mesh.orientation[3] = 1.0;
    mesh.scale = 1.0; mesh.padding = 0.0;
    built.meshes.push_back(mesh);
    auto dev = std::make_shared<StompOptimizer>(start, goal, &robot, &params, &built);
    if (!dev->ok()) { std::printf("FAIL device-built field: %s\n", lastError()); return 1; }
    int32_t dims[3] = {0, 0, 0}, vt = -1;
    std::vector<uint8_t> vox(size_t(60) * 60 * 60);
    if (stomp_engine_get_sdf(dev->getEngine()->get(), dims, &vt, vox.data(), vox.size()) || dims[0] != 60 || vt != STOMP_VOXEL_U8_SQ) {
      std::printf("FAIL get_sdf: %s\n", lastError());
      return 1;
    }
    auto cell = [&](double x, double y, double z) {
      return vox[(size_t(std::lround((x + 0.9) / 0.03)) * 60 + size_t(std::lround((y + 0.9) / 0.03))) * 60 + size_t(std::lround((z + 0.2) / 0.03))];
    };
    if (cell(0.55, 0.0, 0.55) != 0 || cell(-0.5, 0.0, 0.3) != 0 || cell(0.0, 0.6, 1.2) != 144) { std::printf("FAIL field contents\n"); return 1; }
    // inside the mesh cube, and one cell beyond its +x face (the face is at x = -0.352: the next lattice cell is 1 away)
    if (cell(-0.45, 0.6, 1.0) != 0 || cell(-0.45 + 0.15, 0.6, 1.0) == 0) { std::printf("FAIL mesh body in the field\n"); return 1; }
    STOMPStatistics st2;
    if (!dev->optimize(&st2) || !(st2.best_cost <= st2.costs.front())) { std::printf("FAIL optimize on the device-built field\n"); return 1; }
  }
  std::printf("facade ok\n");
  return 0;
}
