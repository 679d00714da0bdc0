// oracle/ref_shim/planning_environment/... — opaque stand-ins: the cost plugin only holds pointers to these.  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_PLANNING_ENV
#define STOMP_REF_SHIM_PLANNING_ENV
#include <tf/transform_listener.h>
namespace bodies { class Body {}; }
namespace planning_environment {
class RobotModels {};
class CollisionModels : public RobotModels {};
class CollisionSpaceMonitor {
 public:
  CollisionModels* getCollisionModels() const { return 0; }
};
}
#endif
