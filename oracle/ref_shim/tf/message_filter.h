// oracle/ref_shim (test infrastructure)
#include <tf/transform_listener.h>
