// oracle/ref_shim/kdl/chainidsolver_recursive_newton_euler.hpp — interface only.  The inverse-dynamics (torque) cost
// term has weight 0 in every shipped configuration (src/stomp_parameters.cpp:56) and is outside the pinned path; the
// stand-in reports zero torques (only ever published as statistics by StompOptimizer::optimize), so the torque term is NOT pinned.  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_KDL_RNE
#define STOMP_REF_SHIM_KDL_RNE
#include <cstdlib>
#include <vector>
#include <kdl/chain.hpp>
#include <kdl/jntarray.hpp>

namespace KDL {
typedef std::vector<Wrench> Wrenches;
class ChainIdSolver {
 public:
  virtual int CartToJnt(const JntArray& q, const JntArray& q_dot, const JntArray& q_dotdot, const Wrenches& f_ext, JntArray& torques) = 0;
  virtual ~ChainIdSolver() {}
};
class ChainIdSolver_RNE : public ChainIdSolver {
 public:
  ChainIdSolver_RNE(const Chain&, Vector) {}
  int CartToJnt(const JntArray&, const JntArray&, const JntArray&, const Wrenches&, JntArray& torques) {
    for (unsigned int i = 0; i < torques.rows(); ++i) torques(i) = 0.0;
    return 0;
  }
};
}  // namespace KDL
#endif
