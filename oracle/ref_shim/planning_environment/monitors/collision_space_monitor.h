// oracle/ref_shim/planning_environment/monitors/collision_space_monitor.h — stand-ins for the ROS arm_navigation stack
// (planning_environment, planning_models, collision_space, geometric_shapes) as far as src/stomp_robot_model.cpp and
// src/stomp_collision_space.cpp name them.  TEST INFRASTRUCTURE.  Functional (filled by oracle/ref_driver.cpp): the list of
// collision-checking links, the environment's collision objects (boxes / cylinders with poses), link padding.  Everything else
// exists so that the unmodified translation units compile; the driver never runs the code paths that need a live robot
// (URDF parsing, robot-body voxelisation, tf).
#ifndef STOMP_REF_SHIM_PLANNING_ENV
#define STOMP_REF_SHIM_PLANNING_ENV
#include <map>
#include <string>
#include <vector>
#include <tf/transform_listener.h>

namespace shapes {
enum ShapeType { UNKNOWN_SHAPE, SPHERE, CYLINDER, BOX, MESH };
class Shape {
 public:
  Shape() : type(UNKNOWN_SHAPE) {}
  virtual ~Shape() {}
  ShapeType type;
};
class Sphere : public Shape { public: explicit Sphere(double r = 0) : radius(r) { type = SPHERE; } double radius; };
class Cylinder : public Shape { public: Cylinder(double r = 0, double l = 0) : length(l), radius(r) { type = CYLINDER; } double length, radius; };
class Box : public Shape { public: Box(double x = 0, double y = 0, double z = 0) { type = BOX; size[0] = x; size[1] = y; size[2] = z; } double size[3]; };
class Mesh : public Shape { public: Mesh() { type = MESH; } };
}  // namespace shapes

namespace bodies {
struct BoundingSphere { btVector3 center; double radius; BoundingSphere() : radius(0) {} };
// geometric_shapes bodies::Body: only the bounding sphere of a padded, posed primitive is functional (what
// StompRobotModel::generateAttachedObjectCollisionPoints reads); ray casting (mesh / robot-body voxelisation) is not.
class Body {
 public:
  explicit Body(const shapes::Shape* s = 0) : shape_(s), padding_(0.0) {}
  virtual ~Body() {}
  void setPadding(double p) { padding_ = p; }
  void setPose(const btTransform& t) { pose_ = t; }
  void computeBoundingSphere(BoundingSphere& s) const {
    s.center = pose_.getOrigin();
    s.radius = 0.0;
    if (!shape_) return;
    if (shape_->type == shapes::SPHERE) s.radius = static_cast<const shapes::Sphere*>(shape_)->radius + padding_;
    if (shape_->type == shapes::BOX) {
      const double* z = static_cast<const shapes::Box*>(shape_)->size;
      double a = z[0] / 2 + padding_, b = z[1] / 2 + padding_, c = z[2] / 2 + padding_;
      s.radius = std::sqrt(a * a + b * b + c * c);
    }
    if (shape_->type == shapes::CYLINDER) {
      const shapes::Cylinder* c = static_cast<const shapes::Cylinder*>(shape_);
      double r = c->radius + padding_, h = c->length / 2 + padding_;
      s.radius = std::sqrt(r * r + h * h);
    }
  }
  bool intersectsRay(const btVector3&, const btVector3&, std::vector<btVector3>* = 0, unsigned int = 0) const { return false; }
 private:
  const shapes::Shape* shape_;
  double padding_;
  btTransform pose_;
};
inline Body* createBodyFromShape(const shapes::Shape* s) { return new Body(s); }
}  // namespace bodies

namespace planning_models {
class KinematicModel {
 public:
  struct Link;
  struct Joint { virtual ~Joint() {} std::string name; };
  struct RevoluteJoint : Joint { bool continuous; double lowLimit, hiLimit; };
  struct PrismaticJoint : Joint { double lowLimit, hiLimit; };
  struct AttachedBody {
    Link* owner;
    std::vector<shapes::Shape*> shapes;
    std::vector<btTransform> attachTrans, globalTrans;
  };
  struct Link {
    std::string name;
    shapes::Shape* shape;
    btTransform globalTransFwd;
    std::vector<AttachedBody*> attachedBodies;
    Link() : shape(0) {}
  };
  void lock() {}
  void unlock() {}
  Joint* getJoint(const std::string&) const { return 0; }
  const Link* getLink(const std::string&) const { return 0; }
  void getJoints(std::vector<const Joint*>&) const {}
  void computeTransforms(const double*) {}
};
class KinematicState {
 public:
  KinematicState() {}
  void setParamsJoint(const std::vector<double>&, const std::string&) {}
  void copyParamsJoint(std::vector<double>&, const std::string&) const {}
  const double* getParams() const { return 0; }
};
}  // namespace planning_models

namespace collision_space {
class EnvironmentObjects {
 public:
  struct NamespaceObjects {
    std::vector<shapes::Shape*> shape;
    std::vector<btTransform> shapePose;
  };
  std::vector<std::string> getNamespaces() const {
    std::vector<std::string> ns;
    for (std::map<std::string, NamespaceObjects>::const_iterator it = objects.begin(); it != objects.end(); ++it) ns.push_back(it->first);
    return ns;
  }
  const NamespaceObjects& getObjects(const std::string& ns) const { return objects.find(ns)->second; }
  std::map<std::string, NamespaceObjects> objects;
};
class EnvironmentModel {
 public:
  EnvironmentModel() : attached_padding(0.0) {}
  void lock() {}
  void unlock() {}
  void updateRobotModel() {}
  const EnvironmentObjects* getObjects() const { return &env_objects; }
  double getCurrentLinkPadding(const std::string&) const { return attached_padding; }
  std::vector<const planning_models::KinematicModel::AttachedBody*> getAttachedBodies() const { return attached; }
  EnvironmentObjects env_objects;
  std::vector<const planning_models::KinematicModel::AttachedBody*> attached;
  double attached_padding;
};
}  // namespace collision_space

namespace planning_environment {
class RobotModels {
 public:
  const std::string& getDescription() const { return description; }
  const std::map<std::string, std::vector<std::string> >& getPlanningGroupLinks() const { return group_links; }
  const std::map<std::string, std::vector<std::string> >& getPlanningGroupJoints() const { return group_joints; }
  const std::vector<std::string>& getGroupLinkUnion() const { return group_link_union; }
  planning_models::KinematicModel* getKinematicModel() const { return const_cast<planning_models::KinematicModel*>(&kmodel); }
  std::string description;
  std::map<std::string, std::vector<std::string> > group_links, group_joints;
  std::vector<std::string> group_link_union;
  planning_models::KinematicModel kmodel;
};
class CollisionModels : public RobotModels {};
class CollisionSpaceMonitor {
 public:
  CollisionModels* getCollisionModels() const { return const_cast<CollisionModels*>(&models); }
  collision_space::EnvironmentModel* getEnvironmentModel() const { return const_cast<collision_space::EnvironmentModel*>(&env); }
  planning_models::KinematicModel* getKinematicModel() const { return models.getKinematicModel(); }
  const planning_models::KinematicState* getRobotState() const { return &state; }
  tf::TransformListener* getTransformListener() const { return const_cast<tf::TransformListener*>(&tfl); }
  void waitForState() const {}
  ros::Time lastJointStateUpdate() const { return ros::Time(); }
  CollisionModels models;
  collision_space::EnvironmentModel env;
  planning_models::KinematicState state;
  tf::TransformListener tfl;
};
}  // namespace planning_environment
#endif
