// oracle/ref_shim/ros_msgs_shim.h — plain structs standing in for the generated ROS message classes that the reference's
// cost-plugin sources name (field names as in the .msg files of the 2010 ROS stacks; only fields the sources touch).
// TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_ROS_MSGS
#define STOMP_REF_SHIM_ROS_MSGS
#include <cstdint>
#include <string>
#include <vector>
#include <ros/ros.h>
#include <boost/shared_ptr.hpp>

namespace std_msgs {
struct Header { std::string frame_id; ros::Time stamp; };
struct ColorRGBA { float r, g, b, a; ColorRGBA() : r(0), g(0), b(0), a(0) {} };
}
namespace geometry_msgs {
struct Point { double x, y, z; Point() : x(0), y(0), z(0) {} };
struct Vector3 { double x, y, z; Vector3() : x(0), y(0), z(0) {} };
struct Quaternion { double x, y, z, w; Quaternion() : x(0), y(0), z(0), w(1) {} };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}
namespace geometric_shapes_msgs {
struct Shape {
  enum { SPHERE = 0, BOX = 1, CYLINDER = 2, MESH = 3 };
  int8_t type;
  std::vector<double> dimensions;
  Shape() : type(0) {}
};
}
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2 };
  std_msgs::Header header;
  std::string ns;
  int32_t id, type, action;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
  Marker() : id(0), type(0), action(0) {}
};
struct MarkerArray { std::vector<Marker> markers; };
}
namespace trajectory_msgs {
struct JointTrajectoryPoint { std::vector<double> positions, velocities, accelerations; ros::Duration time_from_start; };
struct JointTrajectory { std_msgs::Header header; std::vector<std::string> joint_names; std::vector<JointTrajectoryPoint> points; };
}
namespace sensor_msgs {
struct JointState { std_msgs::Header header; std::vector<std::string> name; std::vector<double> position, velocity, effort; };
}
namespace motion_planning_msgs {
struct RobotState { sensor_msgs::JointState joint_state; };
struct OrientationConstraint {
  enum { LINK_FRAME = 0, HEADER_FRAME = 1 };
  std_msgs::Header header;
  std::string link_name;
  int32_t type;
  geometry_msgs::Quaternion orientation;
  double absolute_roll_tolerance, absolute_pitch_tolerance, absolute_yaw_tolerance, weight;
  OrientationConstraint() : type(HEADER_FRAME), absolute_roll_tolerance(0), absolute_pitch_tolerance(0), absolute_yaw_tolerance(0), weight(1) {}
};
struct JointConstraint { std::string joint_name; double position, tolerance_above, tolerance_below, weight; };
struct PositionConstraint {};
struct Constraints {
  std::vector<JointConstraint> joint_constraints;
  std::vector<PositionConstraint> position_constraints;
  std::vector<OrientationConstraint> orientation_constraints;
};
}
namespace mapping_msgs {
struct CollisionMap {};
struct CollisionObjectOperation { enum { ADD = 0, REMOVE = 1 }; int8_t operation; };
struct CollisionObject {};
struct AttachedCollisionObject {};
typedef boost::shared_ptr<const CollisionMap> CollisionMapConstPtr;
typedef boost::shared_ptr<const CollisionObject> CollisionObjectConstPtr;
typedef boost::shared_ptr<const AttachedCollisionObject> AttachedCollisionObjectConstPtr;
}
#endif
