// oracle/ref_shim/stomp_motion_planner/stomp_trajectory.h — shadows the reference header for the one
// translation unit that needs it here, src/stomp_cost.cpp, which only calls getNumPoints() and
// getDiscretization() (src/stomp_cost.cpp:49,57).  Semantics: include/stomp_motion_planner/stomp_trajectory.h:240-258.
#ifndef STOMP_REF_SHIM_TRAJECTORY
#define STOMP_REF_SHIM_TRAJECTORY
#include <Eigen/Core>
namespace stomp_motion_planner {
class StompTrajectory {
 public:
  StompTrajectory(int num_points, double discretization) : num_points_(num_points), discretization_(discretization) {}
  int getNumPoints() const { return num_points_; }
  double getDiscretization() const { return discretization_; }
 private:
  int num_points_;
  double discretization_;
};
}
#endif
