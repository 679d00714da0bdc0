// Host-only check of the request pre / post-processing helpers of the facade (no GPU needed): tests/test_facade_helpers_cpu.py builds and runs it.
#include <cassert>
#include <cstdio>
#include <stomp_motion_planner/stomp_b200_facade.hpp>
using namespace stomp_motion_planner;
int main() {
  assert(std::fabs(shortestAngularDistance(3.0, -3.0) - (2.0 * M_PI - 6.0)) < 1e-15);
  assert(std::fabs(shortestAngularDistance(0.1, 0.4) - 0.3) < 1e-15);
  assert(std::fabs(shortestAngularDistance(-3.0, 3.0) + (2.0 * M_PI - 6.0)) < 1e-15);
  VectorXd start = {3.0, 1.0}, goal = {-3.0, -1.0};
  std::vector<stomp_joint_limit> lim(2);
  lim[0].has_limits = 0; lim[1].has_limits = 1; lim[1].min = -2; lim[1].max = 2;
  fixGoalForWrapAroundJoints(start, goal, lim);
  assert(std::fabs(goal[0] - (3.0 + 2.0 * M_PI - 6.0)) < 1e-15 && goal[1] == -1.0);
  std::vector<VectorXd> traj = {{0.1, 0.2, 0.5}, {0.0, 0.0, 0.0}};
  std::vector<double> t = timeFromStart({0.0, 0.0}, traj, {0.6, 0.0}, 0.05, {2.0});
  // steps 0.1, 0.1, 0.3, 0.1 at 2 rad/s: 0.05 (discretization wins), 0.05, 0.15, 0.05
  const double want[5] = {0.0, 0.05, 0.10, 0.25, 0.30};
  for (int i = 0; i < 5; ++i) assert(std::fabs(t[i] - want[i]) < 1e-15);
  std::puts("request helpers ok");
  return 0;
}
