// oracle/ref_shim/boost/random/variate_generator.hpp (test infrastructure)
#ifndef STOMP_REF_SHIM_BOOST_VARGEN
#define STOMP_REF_SHIM_BOOST_VARGEN
namespace boost {
template <typename Engine, typename Dist>
class variate_generator {
 public:
  variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
  typename Dist::result_type operator()() { return d_(e_); }
 private:
  Engine e_;
  Dist d_;
};
}
#endif
