// oracle/ref_shim/boost/random/mersenne_twister.hpp — boost::mt19937 is the same generator as std::mt19937.
#ifndef STOMP_REF_SHIM_BOOST_MT
#define STOMP_REF_SHIM_BOOST_MT
#include <random>
namespace boost { typedef std::mt19937 mt19937; }
#endif
