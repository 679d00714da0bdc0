// oracle/ref_shim (test infrastructure)
#include <distance_field/distance_field.h>
