// oracle/ref_shim: rosbag is only named by a commented-out block of the reference (src/policy_improvement_loop.cpp:205-243).
#include <ros/ros.h>
