// oracle/ref_shim/kdl/jntarray.hpp — KDL::JntArray as far as include/stomp_motion_planner/stomp_utils.h names it.
#ifndef STOMP_REF_SHIM_KDL_JNTARRAY
#define STOMP_REF_SHIM_KDL_JNTARRAY
#include <vector>
namespace KDL {
class JntArray {
 public:
  JntArray() {}
  explicit JntArray(unsigned int n) : q_(n, 0.0) {}
  void resize(unsigned int n) { q_.assign(n, 0.0); }
  unsigned int rows() const { return (unsigned int)q_.size(); }
  double& operator()(unsigned int i) { return q_[i]; }
  double operator()(unsigned int i) const { return q_[i]; }
 private:
  std::vector<double> q_;
};
}
#endif
