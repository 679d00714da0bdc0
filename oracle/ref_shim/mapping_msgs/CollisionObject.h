// oracle/ref_shim: see ros_msgs_shim.h (test infrastructure).
#include <ros_msgs_shim.h>
