// oracle/ref_shim/ros/ros.h — stand-in for the parts of roscpp the reference's PI^2 sources touch
// (logging macros, assertions, NodeHandle parameter lookup backed by an in-process map).
// TEST INFRASTRUCTURE: lets g++ compile the unmodified reference sources (oracle/Makefile).
#ifndef STOMP_REF_SHIM_ROS_H
#define STOMP_REF_SHIM_ROS_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ref] "); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)

// The reference is built with -DNDEBUG, where ROS_ASSERT compiles to nothing and ROS_ASSERT_FUNC
// (include/stomp_motion_planner/assert.h:43-56) evaluates its argument and ignores the result.
// ROS_ASSERT_ENABLED is therefore left undefined here.
#define ROS_ASSERT(cond) ((void)0)
#define ROS_ASSERT_MSG(cond, ...) ((void)0)
#define ROS_BREAK() std::abort()

namespace XmlRpc {
class XmlRpcValue {
 public:
  enum Type { TypeInvalid, TypeBoolean, TypeInt, TypeDouble, TypeString, TypeArray };
  XmlRpcValue() : type_(TypeInvalid), i_(0), d_(0.0) {}
  XmlRpcValue(int v) : type_(TypeInt), i_(v), d_(0.0) {}
  XmlRpcValue(bool v) : type_(TypeBoolean), i_(v), d_(0.0) {}
  XmlRpcValue(double v) : type_(TypeDouble), i_(0), d_(v) {}
  XmlRpcValue(const std::string& v) : type_(TypeString), i_(0), d_(0.0), s_(v) {}
  XmlRpcValue(const std::vector<double>& v) : type_(TypeArray), i_(0), d_(0.0) {
    for (size_t k = 0; k < v.size(); ++k) a_.push_back(XmlRpcValue(v[k]));
  }
  Type getType() const { return type_; }
  int size() const { return int(a_.size()); }
  XmlRpcValue& operator[](int i) { return a_[size_t(i)]; }
  XmlRpcValue& operator[](const char* k) { return m_[k]; }
  XmlRpcValue& operator[](const std::string& k) { return m_[k]; }
  bool hasMember(const std::string& k) const { return m_.count(k) != 0; }
  operator int&() { return i_; }
  operator double&() { return d_; }
  operator std::string&() { return s_; }
  bool asBool() const { return i_ != 0; }
  int asInt() const { return i_; }
  double asDouble() const { return d_; }
  const std::string& asString() const { return s_; }
 private:
  Type type_;
  int i_;
  double d_;
  std::string s_;
  std::vector<XmlRpcValue> a_;
  std::map<std::string, XmlRpcValue> m_;
};
}  // namespace XmlRpc

namespace ros {

class NodeHandle {
 public:
  typedef std::map<std::string, XmlRpc::XmlRpcValue> Params;
  // every handle of one namespace sees the same parameters, like handles onto one parameter server
  static std::shared_ptr<Params> server(const std::string& ns) {
    static std::map<std::string, std::shared_ptr<Params> > servers;
    std::shared_ptr<Params>& p = servers[ns];
    if (!p) p.reset(new Params);
    return p;
  }
  NodeHandle() : ns_("~"), params_(server("~")) {}
  explicit NodeHandle(const std::string& ns) : ns_(ns), params_(server(ns)) {}
  const std::string& getNamespace() const { return ns_; }

  // test-side population
  void set(const std::string& k, const XmlRpc::XmlRpcValue& v) { (*params_)[k] = v; }
  void erase(const std::string& k) { params_->erase(k); }

  bool getParam(const std::string& k, XmlRpc::XmlRpcValue& v) const {
    Params::const_iterator it = params_->find(k);
    if (it == params_->end()) return false;
    v = it->second; return true;
  }
  bool getParam(const std::string& k, int& v) const {
    Params::const_iterator it = params_->find(k);
    if (it == params_->end() || it->second.getType() != XmlRpc::XmlRpcValue::TypeInt) return false;
    v = it->second.asInt(); return true;
  }
  bool getParam(const std::string& k, double& v) const {
    Params::const_iterator it = params_->find(k);
    if (it == params_->end()) return false;
    if (it->second.getType() == XmlRpc::XmlRpcValue::TypeDouble) { v = it->second.asDouble(); return true; }
    if (it->second.getType() == XmlRpc::XmlRpcValue::TypeInt) { v = it->second.asInt(); return true; }
    return false;
  }
  bool getParam(const std::string& k, bool& v) const {
    Params::const_iterator it = params_->find(k);
    if (it == params_->end() || it->second.getType() != XmlRpc::XmlRpcValue::TypeBoolean) return false;
    v = it->second.asBool(); return true;
  }
  bool getParam(const std::string& k, std::string& v) const {
    Params::const_iterator it = params_->find(k);
    if (it == params_->end() || it->second.getType() != XmlRpc::XmlRpcValue::TypeString) return false;
    v = it->second.asString(); return true;
  }
  template <typename T> bool param(const std::string& k, T& v, const T& dflt) const {
    if (getParam(k, v)) return true;
    v = dflt; return false;
  }
  bool hasParam(const std::string& k) const { return params_->count(k) != 0; }

 private:
  std::string ns_;
  std::shared_ptr<Params> params_;
};


class WallDuration {
 public:
  explicit WallDuration(double s = 0.0) : s_(s) {}
  double toSec() const { return s_; }
  bool sleep() const { return true; }
 private:
  double s_;
};
class WallTime {
 public:
  static WallTime now() { return WallTime(); }
  WallDuration operator-(const WallTime&) const { return WallDuration(0.0); }
};
class Duration {
 public:
  explicit Duration(double s = 0.0) : s_(s) {}
  double toSec() const { return s_; }
  bool sleep() const { return true; }
  Duration operator-(const Duration& o) const { return Duration(s_ - o.s_); }
 private:
  double s_;
};
class Time {
 public:
  Time() {}
  static Time now() { return Time(); }
};
inline bool ok() { return true; }
inline void spinOnce() {}

// A publisher that keeps the last message of each type it was given (read back by oracle/ref_driver.cpp).
class Publisher {
 public:
  template <typename M> void publish(const M& m) const { last<M>() = m; }
  template <typename M> static M& last() { static M m; return m; }
};
class Subscriber {};

}  // namespace ros

#endif
