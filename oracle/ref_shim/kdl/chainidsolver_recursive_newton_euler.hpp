// oracle/ref_shim/kdl/chainidsolver_recursive_newton_euler.hpp — stand-in for KDL::ChainIdSolver_RNE (orocos KDL 1.0):
// recursive Newton-Euler inverse dynamics of a serial chain in body (segment tip) coordinates, after Featherstone, with the
// gravity handed over as a base acceleration.  Third-party algorithm restated (it is not the reference's own code), TEST
// INFRASTRUCTURE.  The physics is checked independently of this file by tests/test_torque_cpu.py (Lagrangian by finite
// differences of the chain's energy).
#ifndef STOMP_REF_SHIM_KDL_RNE
#define STOMP_REF_SHIM_KDL_RNE
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
  ChainIdSolver_RNE(const Chain& chain_, Vector grav)
      : chain(chain_), nj(chain.getNrOfJoints()), ns(chain.getNrOfSegments()), X(ns), S(ns), v(ns), a(ns), f(ns),
        ag(-Twist(grav, Vector::Zero())) {}
  int CartToJnt(const JntArray& q, const JntArray& q_dot, const JntArray& q_dotdot, const Wrenches& f_ext, JntArray& torques) {
    if (q.rows() != nj || q_dot.rows() != nj || q_dotdot.rows() != nj || torques.rows() != nj || f_ext.size() != ns) return -1;
    unsigned int j = 0;
    for (unsigned int i = 0; i < ns; i++) {   // root -> leaf: velocities, accelerations, net wrench of every segment
      double q_, qdot_, qdotdot_;
      if (chain.getSegment(i).getJoint().getType() != Joint::None) {
        q_ = q(j); qdot_ = q_dot(j); qdotdot_ = q_dotdot(j);
        j++;
      } else {
        q_ = qdot_ = qdotdot_ = 0.0;
      }
      X[i] = chain.getSegment(i).pose(q_);
      Twist vj = X[i].M.Inverse(chain.getSegment(i).twist(q_, qdot_));
      S[i] = X[i].M.Inverse(chain.getSegment(i).twist(q_, 1.0));
      if (i == 0) {
        v[i] = vj;
        a[i] = X[i].Inverse(ag) + S[i] * qdotdot_ + v[i] * vj;
      } else {
        v[i] = X[i].Inverse(v[i - 1]) + vj;
        a[i] = X[i].Inverse(a[i - 1]) + S[i] * qdotdot_ + v[i] * vj;
      }
      const RigidBodyInertia& Ii = chain.getSegment(i).getInertia();
      f[i] = Ii * a[i] + v[i] * (Ii * v[i]) - f_ext[i];
    }
    j = nj;
    for (int i = int(ns) - 1; i >= 0; i--) {  // leaf -> root: joint torques, wrench handed to the parent
      if (chain.getSegment(i).getJoint().getType() != Joint::None) torques(--j) = dot(S[i], f[i]);
      if (i != 0) f[i - 1] = f[i - 1] + X[i] * f[i];
    }
    return 0;
  }

 private:
  Chain chain;
  unsigned int nj, ns;
  std::vector<Frame> X;
  std::vector<Twist> S, v, a;
  std::vector<Wrench> f;
  Twist ag;
};
}  // namespace KDL
#endif
