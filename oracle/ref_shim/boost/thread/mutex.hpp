// oracle/ref_shim/boost/thread/mutex.hpp — boost::mutex spelled with std::mutex (test infrastructure).
#ifndef STOMP_REF_SHIM_BOOST_MUTEX
#define STOMP_REF_SHIM_BOOST_MUTEX
#include <mutex>
namespace boost { typedef std::mutex mutex; }
#endif
