// oracle/ref_shim/stomp_motion_planner/stomp_robot_model.h — shadows the reference header of the same name,
// which stomp_utils.h includes without using it for DIFF_RULES; the real header drags in ROS messages,
// planning_environment and KDL trees that the PI^2 translation units never touch.
#ifndef STOMP_REF_SHIM_ROBOT_MODEL
#define STOMP_REF_SHIM_ROBOT_MODEL
#endif
