// oracle/ref_shim/kdl/chain.hpp — KDL::Joint, KDL::Segment, KDL::Chain stand-ins (see kdl/tree.hpp).  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_KDL_CHAIN
#define STOMP_REF_SHIM_KDL_CHAIN
#include <string>
#include <vector>
#include <kdl/frames.hpp>

namespace KDL {

class Joint {
 public:
  enum JointType { RotAxis, RotX, RotY, RotZ, TransAxis, TransX, TransY, TransZ, None };
  Joint() : type_(None) {}
  explicit Joint(const std::string& name, JointType type = None) : name_(name), type_(type) {}
  Joint(const std::string& name, const Vector& origin, const Vector& axis, JointType type)
      : name_(name), type_(type), origin_(origin), axis_(axis / axis.Norm()) {}
  Frame pose(double q) const {
    switch (type_) {
      case RotAxis: return Frame(Rotation::Rot2(axis_, q), origin_);
      case TransAxis: return Frame(origin_ + q * axis_);
      default: return Frame::Identity();
    }
  }
  Vector JointAxis() const { return type_ == None ? Vector::Zero() : axis_; }
  Vector JointOrigin() const { return origin_; }
  const std::string& getName() const { return name_; }
  JointType getType() const { return type_; }

 private:
  std::string name_;
  JointType type_;
  Vector origin_, axis_;
};

class RigidBodyInertia {};

class Segment {
 public:
  Segment() {}
  Segment(const std::string& name, const Joint& joint = Joint(), const Frame& f_tip = Frame::Identity(),
          const RigidBodyInertia& = RigidBodyInertia())
      : name_(name), joint_(joint), f_tip_(joint.pose(0).Inverse() * f_tip) {}
  Frame pose(double q) const { return joint_.pose(q) * f_tip_; }
  const Joint& getJoint() const { return joint_; }
  const std::string& getName() const { return name_; }
  Frame getFrameToTip() const { return joint_.pose(0) * f_tip_; }

 private:
  std::string name_;
  Joint joint_;
  Frame f_tip_;
};

class Chain {
 public:
  Chain() : nj_(0) {}
  void addSegment(const Segment& s) { segments_.push_back(s); if (s.getJoint().getType() != Joint::None) nj_++; }
  unsigned int getNrOfSegments() const { return (unsigned int)segments_.size(); }
  unsigned int getNrOfJoints() const { return nj_; }
  const Segment& getSegment(unsigned int i) const { return segments_[i]; }

 private:
  std::vector<Segment> segments_;
  unsigned int nj_;
};

}  // namespace KDL
#endif
