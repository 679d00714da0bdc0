// oracle/ref_shim/stomp_motion_planner/STOMPStatistics.h — the class rosbuild would generate from the reference's
// msg/STOMPStatistics.msg (same fields, same order).  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_STATISTICS
#define STOMP_REF_SHIM_STATISTICS
#include <cstdint>
#include <vector>
namespace stomp_motion_planner {
struct STOMPStatistics {
  bool success;
  int32_t success_iteration;
  double success_duration;
  int32_t collision_success_iteration;
  double collision_success_duration;
  double best_cost;
  std::vector<double> costs;
  std::vector<double> torques;
  STOMPStatistics() : success(false), success_iteration(0), success_duration(0), collision_success_iteration(0), collision_success_duration(0), best_cost(0) {}
};
}
#endif
