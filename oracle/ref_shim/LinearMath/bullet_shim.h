// oracle/ref_shim/LinearMath/bullet_shim.h — stand-in for the parts of Bullet's LinearMath (as patched by ROS tf, 2010)
// that src/constraint_evaluator.cpp uses: btQuaternion, btMatrix3x3(q), inverse(), operator*, getRPY (= getEulerYPR,
// solution 1).  Third-party semantics restated, TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_BULLET
#define STOMP_REF_SHIM_BULLET
#include <cmath>
typedef double btScalar;
class btVector3 {
 public:
  btVector3() { v_[0] = v_[1] = v_[2] = 0.0; }
  btVector3(double x, double y, double z) { v_[0] = x; v_[1] = y; v_[2] = z; }
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  void setX(double v) { v_[0] = v; }
  void setY(double v) { v_[1] = v; }
  void setZ(double v) { v_[2] = v; }
 private:
  double v_[3];
};
class btQuaternion {
 public:
  btQuaternion() : x_(0), y_(0), z_(0), w_(1) {}
  btQuaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double w() const { return w_; }
  double length2() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
 private:
  double x_, y_, z_, w_;
};
class btMatrix3x3 {
 public:
  btMatrix3x3() { for (int i = 0; i < 9; ++i) m_[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  explicit btMatrix3x3(const btQuaternion& q) { setRotation(q); }
  void setRotation(const btQuaternion& q) {
    double d = q.length2(), s = 2.0 / d;
    double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    m_[0] = 1.0 - (yy + zz); m_[1] = xy - wz; m_[2] = xz + wy;
    m_[3] = xy + wz; m_[4] = 1.0 - (xx + zz); m_[5] = yz - wx;
    m_[6] = xz - wy; m_[7] = yz + wx; m_[8] = 1.0 - (xx + yy);
  }
  double at(int i, int j) const { return m_[i * 3 + j]; }
  btMatrix3x3 inverse() const {  // general 3x3 inverse through cofactors, as Bullet does
    double c0 = m_[4] * m_[8] - m_[5] * m_[7], c1 = m_[5] * m_[6] - m_[3] * m_[8], c2 = m_[3] * m_[7] - m_[4] * m_[6];
    double det = m_[0] * c0 + m_[1] * c1 + m_[2] * c2, s = 1.0 / det;
    btMatrix3x3 r;
    r.m_[0] = c0 * s; r.m_[1] = (m_[2] * m_[7] - m_[1] * m_[8]) * s; r.m_[2] = (m_[1] * m_[5] - m_[2] * m_[4]) * s;
    r.m_[3] = c1 * s; r.m_[4] = (m_[0] * m_[8] - m_[2] * m_[6]) * s; r.m_[5] = (m_[2] * m_[3] - m_[0] * m_[5]) * s;
    r.m_[6] = c2 * s; r.m_[7] = (m_[1] * m_[6] - m_[0] * m_[7]) * s; r.m_[8] = (m_[0] * m_[4] - m_[1] * m_[3]) * s;
    return r;
  }
  btMatrix3x3 operator*(const btMatrix3x3& b) const {
    btMatrix3x3 c;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) c.m_[i * 3 + j] = m_[i * 3] * b.m_[j] + m_[i * 3 + 1] * b.m_[3 + j] + m_[i * 3 + 2] * b.m_[6 + j];
    return c;
  }
  void getEulerYPR(double& yaw, double& pitch, double& roll) const {  // solution 1
    if (std::fabs(m_[6]) >= 1.0) {
      yaw = 0.0;
      double delta = std::atan2(m_[7], m_[8]);
      pitch = m_[6] < 0.0 ? M_PI / 2.0 : -M_PI / 2.0;
      roll = delta;
    } else {
      pitch = -std::asin(m_[6]);
      double cp = std::cos(pitch);
      roll = std::atan2(m_[7] / cp, m_[8] / cp);
      yaw = std::atan2(m_[3] / cp, m_[0] / cp);
    }
  }
  void getRPY(double& roll, double& pitch, double& yaw) const { getEulerYPR(yaw, pitch, roll); }
 private:
  double m_[9];
};
class btTransform {
 public:
  btTransform() {}
  btTransform(const btQuaternion& q, const btVector3& o) : q_(q), o_(o) {}
  const btVector3& getOrigin() const { return o_; }
  const btQuaternion& getRotation() const { return q_; }
  void setOrigin(const btVector3& o) { o_ = o; }
  void setRotation(const btQuaternion& q) { q_ = q; }
 private:
  btQuaternion q_;
  btVector3 o_;
};
#endif
