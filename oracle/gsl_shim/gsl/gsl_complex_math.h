/* Part of oracle/gsl_shim: forwards to the single-header GSL-compatible shim (TEST INFRASTRUCTURE ONLY). */
#ifndef ORACLE_SHIM_GSL_COMPLEX_MATH_H
#define ORACLE_SHIM_GSL_COMPLEX_MATH_H
#include "gsl_shim.h"
#endif
