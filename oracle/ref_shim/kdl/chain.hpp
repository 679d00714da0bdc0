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
  Twist twist(double qdot) const {
    switch (type_) {
      case RotAxis: return Twist(Vector::Zero(), axis_ * qdot);
      case TransAxis: return Twist(axis_ * qdot, Vector::Zero());
      default: return Twist::Zero();
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

// KDL::RigidBodyInertia(m, oc, Ic): mass, centre of gravity oc and rotational inertia Ic ABOUT THE COG, both expressed in
// the frame the inertia is used in (the segment's tip frame); stored as m, h = m * oc and I about the frame origin
// (I = Ic - m * hat(oc)^2).  Ic = (ixx, iyy, izz, ixy, ixz, iyz).
class RigidBodyInertia {
 public:
  double m;
  Vector h;
  double I[9];
  RigidBodyInertia() : m(0.0) { for (int i = 0; i < 9; ++i) I[i] = 0.0; }
  RigidBodyInertia(double m_, const Vector& oc, const double Ic[6]) : m(m_), h(m_ * oc) {
    const double c0 = oc(0), c1 = oc(1), c2 = oc(2), cc = c0 * c0 + c1 * c1 + c2 * c2;
    I[0] = Ic[0] + m * (cc - c0 * c0); I[1] = Ic[3] - m * c0 * c1;         I[2] = Ic[4] - m * c0 * c2;
    I[3] = I[1];                       I[4] = Ic[1] + m * (cc - c1 * c1); I[5] = Ic[5] - m * c1 * c2;
    I[6] = I[2];                       I[7] = I[5];                        I[8] = Ic[2] + m * (cc - c2 * c2);
  }
  Vector rotate(const Vector& w) const {
    return Vector(I[0] * w(0) + I[1] * w(1) + I[2] * w(2), I[3] * w(0) + I[4] * w(1) + I[5] * w(2), I[6] * w(0) + I[7] * w(1) + I[8] * w(2));
  }
};
inline Wrench operator*(const RigidBodyInertia& I, const Twist& t) { return Wrench(I.m * t.vel - I.h * t.rot, I.rotate(t.rot) + I.h * t.vel); }

class Segment {
 public:
  Segment() {}
  Segment(const std::string& name, const Joint& joint = Joint(), const Frame& f_tip = Frame::Identity(),
          const RigidBodyInertia& I = RigidBodyInertia())
      : name_(name), joint_(joint), I_(I), f_tip_(joint.pose(0).Inverse() * f_tip) {}
  Frame pose(double q) const { return joint_.pose(q) * f_tip_; }
  Twist twist(double q, double qdot) const { return joint_.twist(qdot).RefPoint(joint_.pose(q).M * f_tip_.p); }
  const RigidBodyInertia& getInertia() const { return I_; }
  const Joint& getJoint() const { return joint_; }
  const std::string& getName() const { return name_; }
  Frame getFrameToTip() const { return joint_.pose(0) * f_tip_; }

 private:
  std::string name_;
  Joint joint_;
  RigidBodyInertia I_;
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
