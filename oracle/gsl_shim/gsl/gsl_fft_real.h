/* forwarding header of the local GSL-compatible shim (oracle/gsl_shim/gsl/gsl_shim.h); declares nothing of its own */
#include "gsl_shim.h"
