// oracle/ref_shim (test infrastructure)
#ifndef STOMP_REF_SHIM_MF
#define STOMP_REF_SHIM_MF
namespace message_filters { template <typename M> class Subscriber {}; }
#endif
