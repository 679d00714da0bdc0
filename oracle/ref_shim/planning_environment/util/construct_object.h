// oracle/ref_shim/planning_environment/util/construct_object.h — planning_environment::constructObjectMsg: shapes::Shape ->
// geometric_shapes_msgs::Shape {type, dimensions} (sphere: r | box: x, y, z | cylinder: r, length).  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_CONSTRUCT_OBJECT
#define STOMP_REF_SHIM_CONSTRUCT_OBJECT
#include <planning_environment/monitors/collision_space_monitor.h>
#include <ros_msgs_shim.h>
namespace planning_environment {
inline bool constructObjectMsg(const shapes::Shape* shape, geometric_shapes_msgs::Shape& obj) {
  obj.dimensions.clear();
  switch (shape->type) {
    case shapes::SPHERE:
      obj.type = geometric_shapes_msgs::Shape::SPHERE;
      obj.dimensions.push_back(static_cast<const shapes::Sphere*>(shape)->radius);
      return true;
    case shapes::BOX: {
      obj.type = geometric_shapes_msgs::Shape::BOX;
      const double* sz = static_cast<const shapes::Box*>(shape)->size;
      obj.dimensions.assign(sz, sz + 3);
      return true;
    }
    case shapes::CYLINDER:
      obj.type = geometric_shapes_msgs::Shape::CYLINDER;
      obj.dimensions.push_back(static_cast<const shapes::Cylinder*>(shape)->radius);
      obj.dimensions.push_back(static_cast<const shapes::Cylinder*>(shape)->length);
      return true;
    default:
      return false;
  }
}
}
#endif
