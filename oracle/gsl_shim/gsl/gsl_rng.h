/* see gsl_math.h — shim for the reference oracle build only */
#include "gsl_math.h"
