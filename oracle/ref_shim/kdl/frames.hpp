// oracle/ref_shim/kdl/frames.hpp — stand-in for the orocos-KDL geometry primitives the reference's cost plugin uses
// (KDL::Vector / Rotation / Frame / Twist / Wrench).  TEST INFRASTRUCTURE: lets g++ compile the unmodified
// reference sources (oracle/Makefile).  orocos KDL is not vendored by the reference and not installed here; the
// semantics below restate its published behaviour (Frame composition, Rotation::Rot2 = Rodrigues formula of a unit axis,
// Rotation::GetQuaternion).  They are third-party semantics, not the reference's own code.
#ifndef STOMP_REF_SHIM_KDL_FRAMES
#define STOMP_REF_SHIM_KDL_FRAMES
#include <cmath>

namespace KDL {

const double epsilon = 0.000001;

class Vector {
 public:
  double data[3];
  Vector() { data[0] = data[1] = data[2] = 0.0; }
  Vector(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
  double operator()(int i) const { return data[i]; }
  double& operator()(int i) { return data[i]; }
  double operator[](int i) const { return data[i]; }
  double& operator[](int i) { return data[i]; }
  double x() const { return data[0]; }
  double y() const { return data[1]; }
  double z() const { return data[2]; }
  void x(double v) { data[0] = v; }
  void y(double v) { data[1] = v; }
  void z(double v) { data[2] = v; }
  Vector& operator+=(const Vector& o) { data[0] += o.data[0]; data[1] += o.data[1]; data[2] += o.data[2]; return *this; }
  Vector& operator-=(const Vector& o) { data[0] -= o.data[0]; data[1] -= o.data[1]; data[2] -= o.data[2]; return *this; }
  double Norm() const { return std::sqrt(data[0] * data[0] + data[1] * data[1] + data[2] * data[2]); }
  static Vector Zero() { return Vector(); }
};
inline Vector operator+(const Vector& a, const Vector& b) { return Vector(a.data[0] + b.data[0], a.data[1] + b.data[1], a.data[2] + b.data[2]); }
inline Vector operator-(const Vector& a, const Vector& b) { return Vector(a.data[0] - b.data[0], a.data[1] - b.data[1], a.data[2] - b.data[2]); }
inline Vector operator-(const Vector& a) { return Vector(-a.data[0], -a.data[1], -a.data[2]); }
inline Vector operator*(const Vector& a, double s) { return Vector(a.data[0] * s, a.data[1] * s, a.data[2] * s); }
inline Vector operator*(double s, const Vector& a) { return Vector(a.data[0] * s, a.data[1] * s, a.data[2] * s); }
inline Vector operator/(const Vector& a, double s) { return Vector(a.data[0] / s, a.data[1] / s, a.data[2] / s); }
inline Vector operator*(const Vector& a, const Vector& b) {  // cross product, as in KDL
  return Vector(a.data[1] * b.data[2] - a.data[2] * b.data[1], a.data[2] * b.data[0] - a.data[0] * b.data[2],
                a.data[0] * b.data[1] - a.data[1] * b.data[0]);
}
inline double dot(const Vector& a, const Vector& b) { return a.data[0] * b.data[0] + a.data[1] * b.data[1] + a.data[2] * b.data[2]; }
inline void SetToZero(Vector& v) { v = Vector::Zero(); }

class Twist;
class Wrench;
class Rotation {
 public:
  double data[9];  // row-major
  Rotation() { for (int i = 0; i < 9; ++i) data[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  Rotation(double Xx, double Yx, double Zx, double Xy, double Yy, double Zy, double Xz, double Yz, double Zz) {
    data[0] = Xx; data[1] = Yx; data[2] = Zx; data[3] = Xy; data[4] = Yy; data[5] = Zy; data[6] = Xz; data[7] = Yz; data[8] = Zz;
  }
  double operator()(int i, int j) const { return data[i * 3 + j]; }
  double& operator()(int i, int j) { return data[i * 3 + j]; }
  static Rotation Identity() { return Rotation(); }
  Rotation Inverse() const { return Rotation(data[0], data[3], data[6], data[1], data[4], data[7], data[2], data[5], data[8]); }
  Vector operator*(const Vector& v) const {
    return Vector(data[0] * v.data[0] + data[1] * v.data[1] + data[2] * v.data[2],
                  data[3] * v.data[0] + data[4] * v.data[1] + data[5] * v.data[2],
                  data[6] * v.data[0] + data[7] * v.data[1] + data[8] * v.data[2]);
  }
  Vector Inverse(const Vector& v) const { return Inverse() * v; }
  inline Twist Inverse(const Twist& t) const;
  inline Twist operator*(const Twist& t) const;
  // rotation of `angle` about a unit-length axis
  static Rotation Rot2(const Vector& v, double angle) {
    double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
    double m_vt_0 = vt * v(0), m_vt_1 = vt * v(1), m_vt_2 = vt * v(2);
    double m_st_0 = v(0) * st, m_st_1 = v(1) * st, m_st_2 = v(2) * st;
    double m_vt_0_1 = m_vt_0 * v(1), m_vt_0_2 = m_vt_0 * v(2), m_vt_1_2 = m_vt_1 * v(2);
    return Rotation(ct + m_vt_0 * v(0), -m_st_2 + m_vt_0_1, m_st_1 + m_vt_0_2,
                    m_st_2 + m_vt_0_1, ct + m_vt_1 * v(1), -m_st_0 + m_vt_1_2,
                    -m_st_1 + m_vt_0_2, m_st_0 + m_vt_1_2, ct + m_vt_2 * v(2));
  }
  static Rotation Rot(const Vector& axis, double angle) {
    double n = axis.Norm();
    return n < epsilon ? Identity() : Rot2(axis / n, angle);
  }
  static Rotation RotX(double a) { double c = std::cos(a), s = std::sin(a); return Rotation(1, 0, 0, 0, c, -s, 0, s, c); }
  static Rotation RotY(double a) { double c = std::cos(a), s = std::sin(a); return Rotation(c, 0, s, 0, 1, 0, -s, 0, c); }
  static Rotation RotZ(double a) { double c = std::cos(a), s = std::sin(a); return Rotation(c, -s, 0, s, c, 0, 0, 0, 1); }
  static Rotation Quaternion(double x, double y, double z, double w) {
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    return Rotation(w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
                    2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x,
                    2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2);
  }
  void GetQuaternion(double& x, double& y, double& z, double& w) const {
    const Rotation& m = *this;
    double trace = m(0, 0) + m(1, 1) + m(2, 2) + 1.0;
    if (trace > epsilon) {
      double s = 0.5 / std::sqrt(trace);
      w = 0.25 / s;
      x = (m(2, 1) - m(1, 2)) * s;
      y = (m(0, 2) - m(2, 0)) * s;
      z = (m(1, 0) - m(0, 1)) * s;
    } else if (m(0, 0) > m(1, 1) && m(0, 0) > m(2, 2)) {
      double s = 2.0 * std::sqrt(1.0 + m(0, 0) - m(1, 1) - m(2, 2));
      w = (m(2, 1) - m(1, 2)) / s; x = 0.25 * s; y = (m(0, 1) + m(1, 0)) / s; z = (m(0, 2) + m(2, 0)) / s;
    } else if (m(1, 1) > m(2, 2)) {
      double s = 2.0 * std::sqrt(1.0 + m(1, 1) - m(0, 0) - m(2, 2));
      w = (m(0, 2) - m(2, 0)) / s; x = (m(0, 1) + m(1, 0)) / s; y = 0.25 * s; z = (m(1, 2) + m(2, 1)) / s;
    } else {
      double s = 2.0 * std::sqrt(1.0 + m(2, 2) - m(0, 0) - m(1, 1));
      w = (m(1, 0) - m(0, 1)) / s; x = (m(0, 2) + m(2, 0)) / s; y = (m(1, 2) + m(2, 1)) / s; z = 0.25 * s;
    }
  }
};
inline Rotation operator*(const Rotation& a, const Rotation& b) {
  Rotation c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.data[i * 3 + j] = a.data[i * 3] * b.data[j] + a.data[i * 3 + 1] * b.data[3 + j] + a.data[i * 3 + 2] * b.data[6 + j];
  return c;
}

class Frame {
 public:
  Vector p;
  Rotation M;
  Frame() {}
  Frame(const Rotation& R, const Vector& V) : p(V), M(R) {}
  explicit Frame(const Vector& V) : p(V) {}
  explicit Frame(const Rotation& R) : M(R) {}
  static Frame Identity() { return Frame(); }
  Frame Inverse() const { Rotation Mi = M.Inverse(); return Frame(Mi, -(Mi * p)); }
  Vector operator*(const Vector& v) const { return M * v + p; }
  inline Twist operator*(const Twist& t) const;
  inline Twist Inverse(const Twist& t) const;
  inline Wrench operator*(const Wrench& w) const;
};
inline Frame operator*(const Frame& a, const Frame& b) { return Frame(a.M * b.M, a.M * b.p + a.p); }

class Twist {
 public:
  Vector vel, rot;
  Twist() {}
  Twist(const Vector& v, const Vector& r) : vel(v), rot(r) {}
  static Twist Zero() { return Twist(); }
  // the same motion seen at a reference point displaced by v_base_AB
  Twist RefPoint(const Vector& v_base_AB) const { return Twist(vel + rot * v_base_AB, rot); }
};
class Wrench {
 public:
  Vector force, torque;
  Wrench() {}
  Wrench(const Vector& f, const Vector& t) : force(f), torque(t) {}
  static Wrench Zero() { return Wrench(); }  // static, as in KDL: `wrenches[i].Zero()` in the reference is a no-op on the element
};
inline Twist operator+(const Twist& a, const Twist& b) { return Twist(a.vel + b.vel, a.rot + b.rot); }
inline Twist operator-(const Twist& a) { return Twist(-a.vel, -a.rot); }
inline Twist operator*(const Twist& a, double s) { return Twist(a.vel * s, a.rot * s); }
inline Twist operator*(double s, const Twist& a) { return Twist(a.vel * s, a.rot * s); }
// spatial cross products (KDL frames.inl)
inline Twist operator*(const Twist& lhs, const Twist& rhs) { return Twist(lhs.rot * rhs.vel + lhs.vel * rhs.rot, lhs.rot * rhs.rot); }
inline Wrench operator*(const Twist& lhs, const Wrench& rhs) { return Wrench(lhs.rot * rhs.force, lhs.rot * rhs.torque + lhs.vel * rhs.force); }
inline Wrench operator+(const Wrench& a, const Wrench& b) { return Wrench(a.force + b.force, a.torque + b.torque); }
inline Wrench operator-(const Wrench& a, const Wrench& b) { return Wrench(a.force - b.force, a.torque - b.torque); }
inline double dot(const Twist& t, const Wrench& w) { return dot(t.vel, w.force) + dot(t.rot, w.torque); }

inline Twist Rotation::Inverse(const Twist& t) const { return Twist(Inverse(t.vel), Inverse(t.rot)); }
inline Twist Rotation::operator*(const Twist& t) const { return Twist((*this) * t.vel, (*this) * t.rot); }
inline Twist Frame::operator*(const Twist& t) const { Vector r = M * t.rot; return Twist(M * t.vel + p * r, r); }
inline Twist Frame::Inverse(const Twist& t) const { return Twist(M.Inverse(t.vel - p * t.rot), M.Inverse(t.rot)); }
inline Wrench Frame::operator*(const Wrench& w) const { Vector f = M * w.force; return Wrench(f, M * w.torque + p * f); }

}  // namespace KDL
#endif
