// oracle/ref_shim/boost/shared_ptr.hpp — boost::shared_ptr spelled with std::shared_ptr (test infrastructure).
#ifndef STOMP_REF_SHIM_BOOST_SHARED_PTR
#define STOMP_REF_SHIM_BOOST_SHARED_PTR
#include <memory>
namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
}
#endif
