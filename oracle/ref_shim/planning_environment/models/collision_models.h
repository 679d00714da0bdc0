// oracle/ref_shim (test infrastructure)
#include <planning_environment/monitors/collision_space_monitor.h>
