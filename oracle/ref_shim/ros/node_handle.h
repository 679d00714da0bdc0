// oracle/ref_shim: see ros/ros.h in this directory (test infrastructure).
#include <ros/ros.h>
