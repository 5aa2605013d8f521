/* oracle/ref_stubs/feature/feature.h — TEST INFRASTRUCTURE ONLY.
 * The reference's dereverberation/dereverberation.h includes "feature/feature.h" but uses nothing from it; the real header
 * pulls in gsl_rng / libsndfile / FFTW, none of which exist here.  This empty stand-in is found first on the include path
 * when (and only when) oracle/Makefile compiles dereverberation.cc. */
#ifndef ORACLE_REF_STUB_FEATURE_H
#define ORACLE_REF_STUB_FEATURE_H
#endif
