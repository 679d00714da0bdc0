// oracle/ref_shim/boost/random/normal_distribution.hpp — a unit normal on top of std::normal_distribution.
// The sequence of draws differs from Boost's Box-Muller; parity tests read the noise the compiled
// reference actually drew (Rollout::noise_) and inject exactly that into the engine, so the draw
// algorithm is immaterial (BASELINE north_star: "the reference's noise is fed in through a host-injection mode").
#ifndef STOMP_REF_SHIM_BOOST_NORMAL
#define STOMP_REF_SHIM_BOOST_NORMAL
#include <random>
namespace boost {
template <typename Real = double>
class normal_distribution {
 public:
  typedef Real result_type;
  typedef Real input_type;
  explicit normal_distribution(Real mean = 0, Real sigma = 1) : d_(mean, sigma) {}
  template <typename Engine> Real operator()(Engine& e) { return d_(e); }
 private:
  std::normal_distribution<Real> d_;
};
}
#endif
