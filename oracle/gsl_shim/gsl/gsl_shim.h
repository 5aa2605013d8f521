/*
 * gsl_shim.h — a minimal, independently written, header-only stand-in for the
 * slice of the GNU Scientific Library API that btk2.0's hot-path sources use.
 *
 * TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own C++
 * (stream/stream.cc, modulated/modulated.cc, beamformer/beamformer.cc,
 * postfilter/postfilter.cc) can be compiled UNMODIFIED from /root/reference
 * into oracle/_ref/ and used as the parity oracle and CPU baseline
 * (GSL itself is not installed and there is no network).
 * Nothing under distant_speech_recognition_b200/ may include or link this.
 *
 * GSL is mathematically standard here (dense vectors/matrices, complex
 * arithmetic, BLAS level 1-3, radix-2 DFT, sinc): any correct double-precision
 * implementation agrees with the real library to ~1e-13, far below the 1e-4
 * parity budget.  Semantics follow the public GSL documentation:
 *   - vectors: {size, stride, data, block, owner}; matrices row-major with tda
 *   - gsl_fft_complex_radix2_forward : X[k] = sum x[n] exp(-2 pi i nk/N)
 *     gsl_fft_complex_radix2_backward: X[k] = sum x[n] exp(+2 pi i nk/N) (unscaled)
 *     gsl_fft_complex_radix2_inverse : backward scaled by 1/N
 *   - gsl_blas_zdotc(x,y) = sum conj(x_i) y_i
 *   - gsl_sf_sinc(x) = sin(pi x)/(pi x)
 */
#ifndef ORACLE_GSL_SHIM_H
#define ORACLE_GSL_SHIM_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ complex */
typedef struct { double dat[2]; } gsl_complex;
typedef struct { float dat[2]; } gsl_complex_float;

#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
#define GSL_SET_COMPLEX(zp, x, y) do { (zp)->dat[0] = (x); (zp)->dat[1] = (y); } while (0)
#define GSL_SET_REAL(zp, x) do { (zp)->dat[0] = (x); } while (0)
#define GSL_SET_IMAG(zp, y) do { (zp)->dat[1] = (y); } while (0)
#define GSL_SUCCESS 0
#define GSL_DBL_EPSILON 2.2204460492503131e-16

static inline gsl_complex gsl_complex_rect(double x, double y) { gsl_complex z; z.dat[0] = x; z.dat[1] = y; return z; }
static inline gsl_complex gsl_complex_polar(double r, double th) { return gsl_complex_rect(r * cos(th), r * sin(th)); }
static inline gsl_complex gsl_complex_add(gsl_complex a, gsl_complex b) { return gsl_complex_rect(a.dat[0] + b.dat[0], a.dat[1] + b.dat[1]); }
static inline gsl_complex gsl_complex_sub(gsl_complex a, gsl_complex b) { return gsl_complex_rect(a.dat[0] - b.dat[0], a.dat[1] - b.dat[1]); }
static inline gsl_complex gsl_complex_mul(gsl_complex a, gsl_complex b) {
  return gsl_complex_rect(a.dat[0] * b.dat[0] - a.dat[1] * b.dat[1], a.dat[0] * b.dat[1] + a.dat[1] * b.dat[0]);
}
static inline gsl_complex gsl_complex_div(gsl_complex a, gsl_complex b) {
  /* scaled division (avoids overflow in |b|^2) */
  double s = 1.0 / hypot(b.dat[0], b.dat[1]);
  double sbr = s * b.dat[0], sbi = s * b.dat[1];
  return gsl_complex_rect((a.dat[0] * sbr + a.dat[1] * sbi) * s, (a.dat[1] * sbr - a.dat[0] * sbi) * s);
}
static inline gsl_complex gsl_complex_add_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] + x, a.dat[1]); }
static inline gsl_complex gsl_complex_sub_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] - x, a.dat[1]); }
static inline gsl_complex gsl_complex_mul_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] * x, a.dat[1] * x); }
static inline gsl_complex gsl_complex_div_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] / x, a.dat[1] / x); }
static inline gsl_complex gsl_complex_conjugate(gsl_complex a) { return gsl_complex_rect(a.dat[0], -a.dat[1]); }
static inline gsl_complex gsl_complex_negative(gsl_complex a) { return gsl_complex_rect(-a.dat[0], -a.dat[1]); }
static inline gsl_complex gsl_complex_inverse(gsl_complex a) {
  double s = 1.0 / hypot(a.dat[0], a.dat[1]);
  return gsl_complex_rect((a.dat[0] * s) * s, -(a.dat[1] * s) * s);
}
static inline double gsl_complex_abs(gsl_complex a) { return hypot(a.dat[0], a.dat[1]); }
static inline double gsl_complex_abs2(gsl_complex a) { return a.dat[0] * a.dat[0] + a.dat[1] * a.dat[1]; }
static inline double gsl_complex_arg(gsl_complex a) { return (a.dat[0] == 0.0 && a.dat[1] == 0.0) ? 0.0 : atan2(a.dat[1], a.dat[0]); }
static inline gsl_complex gsl_complex_exp(gsl_complex a) { return gsl_complex_polar(exp(a.dat[0]), a.dat[1]); }
static inline gsl_complex gsl_complex_sqrt(gsl_complex a) {
  double r = gsl_complex_abs(a), th = gsl_complex_arg(a);
  return gsl_complex_polar(sqrt(r), 0.5 * th);
}
static inline gsl_complex gsl_complex_sqrt_real(double x) {
  return x >= 0 ? gsl_complex_rect(sqrt(x), 0.0) : gsl_complex_rect(0.0, sqrt(-x));
}
static inline gsl_complex gsl_complex_log(gsl_complex a) { return gsl_complex_rect(log(gsl_complex_abs(a)), gsl_complex_arg(a)); }
static inline gsl_complex gsl_complex_pow_real(gsl_complex a, double b) {
  if (a.dat[0] == 0 && a.dat[1] == 0) return gsl_complex_rect(b == 0 ? 1.0 : 0.0, 0.0);
  return gsl_complex_polar(pow(gsl_complex_abs(a), b), gsl_complex_arg(a) * b);
}

/* ------------------------------------------------------------------ blocks, vectors, matrices */
typedef struct { size_t size; double* data; } gsl_block;
typedef struct { size_t size; double* data; } gsl_block_complex;
typedef struct { size_t size; float* data; } gsl_block_float;
typedef struct { size_t size; short* data; } gsl_block_short;
typedef struct { size_t size; char* data; } gsl_block_char;

#define SHIM_DEFINE_REAL_VECTOR(SUFFIX, T, BLOCK)                                                              \
  typedef struct { size_t size; size_t stride; T* data; BLOCK* block; int owner; } gsl_vector##SUFFIX;         \
  static inline gsl_vector##SUFFIX* gsl_vector##SUFFIX##_calloc(size_t n) {                                    \
    gsl_vector##SUFFIX* v = (gsl_vector##SUFFIX*)malloc(sizeof(gsl_vector##SUFFIX));                          \
    BLOCK* b = (BLOCK*)malloc(sizeof(BLOCK));                                                                  \
    b->size = n; b->data = (T*)calloc(n ? n : 1, sizeof(T));                                                   \
    v->size = n; v->stride = 1; v->data = b->data; v->block = b; v->owner = 1; return v; }                     \
  static inline gsl_vector##SUFFIX* gsl_vector##SUFFIX##_alloc(size_t n) { return gsl_vector##SUFFIX##_calloc(n); } \
  static inline void gsl_vector##SUFFIX##_free(gsl_vector##SUFFIX* v) {                                        \
    if (!v) return; if (v->owner && v->block) { free(v->block->data); free(v->block); } free(v); }             \
  static inline T gsl_vector##SUFFIX##_get(const gsl_vector##SUFFIX* v, size_t i) { return v->data[i * v->stride]; } \
  static inline void gsl_vector##SUFFIX##_set(gsl_vector##SUFFIX* v, size_t i, T x) { v->data[i * v->stride] = x; } \
  static inline T* gsl_vector##SUFFIX##_ptr(gsl_vector##SUFFIX* v, size_t i) { return v->data + i * v->stride; } \
  static inline void gsl_vector##SUFFIX##_set_zero(gsl_vector##SUFFIX* v) {                                   \
    for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = (T)0; }                                      \
  static inline void gsl_vector##SUFFIX##_set_all(gsl_vector##SUFFIX* v, T x) {                               \
    for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = x; }                                         \
  static inline int gsl_vector##SUFFIX##_memcpy(gsl_vector##SUFFIX* d, const gsl_vector##SUFFIX* s) {         \
    for (size_t i = 0; i < s->size; i++) d->data[i * d->stride] = s->data[i * s->stride]; return 0; }          \
  static inline int gsl_vector##SUFFIX##_scale(gsl_vector##SUFFIX* v, const double x) {                       \
    for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = (T)(v->data[i * v->stride] * x); return 0; } \
  static inline int gsl_vector##SUFFIX##_add(gsl_vector##SUFFIX* a, const gsl_vector##SUFFIX* b) {            \
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] += b->data[i * b->stride]; return 0; }         \
  static inline int gsl_vector##SUFFIX##_sub(gsl_vector##SUFFIX* a, const gsl_vector##SUFFIX* b) {            \
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] -= b->data[i * b->stride]; return 0; }         \
  static inline int gsl_vector##SUFFIX##_add_constant(gsl_vector##SUFFIX* a, const double x) {                \
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] = (T)(a->data[i * a->stride] + x); return 0; } \
  static inline int gsl_vector##SUFFIX##_fwrite(FILE* fp, const gsl_vector##SUFFIX* v) {                      \
    for (size_t i = 0; i < v->size; i++) if (fwrite(v->data + i * v->stride, sizeof(T), 1, fp) != 1) return 1; \
    return 0; }                                                                                                \
  static inline int gsl_vector##SUFFIX##_fread(FILE* fp, gsl_vector##SUFFIX* v) {                             \
    for (size_t i = 0; i < v->size; i++) if (fread(v->data + i * v->stride, sizeof(T), 1, fp) != 1) return 1;  \
    return 0; }

SHIM_DEFINE_REAL_VECTOR(, double, gsl_block)
SHIM_DEFINE_REAL_VECTOR(_float, float, gsl_block_float)
SHIM_DEFINE_REAL_VECTOR(_short, short, gsl_block_short)
SHIM_DEFINE_REAL_VECTOR(_char, char, gsl_block_char)

static inline double gsl_vector_max(const gsl_vector* v) {
  double m = v->data[0]; for (size_t i = 1; i < v->size; i++) if (v->data[i * v->stride] > m) m = v->data[i * v->stride]; return m; }
static inline double gsl_vector_min(const gsl_vector* v) {
  double m = v->data[0]; for (size_t i = 1; i < v->size; i++) if (v->data[i * v->stride] < m) m = v->data[i * v->stride]; return m; }

/* complex vector: interleaved (re,im) doubles */
typedef struct { size_t size; size_t stride; double* data; gsl_block_complex* block; int owner; } gsl_vector_complex;
static inline gsl_vector_complex* gsl_vector_complex_calloc(size_t n) {
  gsl_vector_complex* v = (gsl_vector_complex*)malloc(sizeof(gsl_vector_complex));
  gsl_block_complex* b = (gsl_block_complex*)malloc(sizeof(gsl_block_complex));
  b->size = n; b->data = (double*)calloc(2 * (n ? n : 1), sizeof(double));
  v->size = n; v->stride = 1; v->data = b->data; v->block = b; v->owner = 1; return v; }
static inline gsl_vector_complex* gsl_vector_complex_alloc(size_t n) { return gsl_vector_complex_calloc(n); }
static inline void gsl_vector_complex_free(gsl_vector_complex* v) {
  if (!v) return; if (v->owner && v->block) { free(v->block->data); free(v->block); } free(v); }
static inline gsl_complex gsl_vector_complex_get(const gsl_vector_complex* v, size_t i) {
  return gsl_complex_rect(v->data[2 * i * v->stride], v->data[2 * i * v->stride + 1]); }
static inline void gsl_vector_complex_set(gsl_vector_complex* v, size_t i, gsl_complex z) {
  v->data[2 * i * v->stride] = z.dat[0]; v->data[2 * i * v->stride + 1] = z.dat[1]; }
static inline gsl_complex* gsl_vector_complex_ptr(gsl_vector_complex* v, size_t i) { return (gsl_complex*)(v->data + 2 * i * v->stride); }
static inline void gsl_vector_complex_set_zero(gsl_vector_complex* v) {
  for (size_t i = 0; i < v->size; i++) { v->data[2 * i * v->stride] = 0; v->data[2 * i * v->stride + 1] = 0; } }
static inline void gsl_vector_complex_set_all(gsl_vector_complex* v, gsl_complex z) {
  for (size_t i = 0; i < v->size; i++) gsl_vector_complex_set(v, i, z); }
static inline int gsl_vector_complex_memcpy(gsl_vector_complex* d, const gsl_vector_complex* s) {
  for (size_t i = 0; i < s->size; i++) gsl_vector_complex_set(d, i, gsl_vector_complex_get(s, i)); return 0; }
static inline int gsl_vector_complex_sub(gsl_vector_complex* a, const gsl_vector_complex* b) {
  for (size_t i = 0; i < a->size; i++) gsl_vector_complex_set(a, i, gsl_complex_sub(gsl_vector_complex_get(a, i), gsl_vector_complex_get(b, i))); return 0; }
static inline int gsl_vector_complex_add(gsl_vector_complex* a, const gsl_vector_complex* b) {
  for (size_t i = 0; i < a->size; i++) gsl_vector_complex_set(a, i, gsl_complex_add(gsl_vector_complex_get(a, i), gsl_vector_complex_get(b, i))); return 0; }
static inline int gsl_vector_complex_scale(gsl_vector_complex* a, const gsl_complex x) {
  for (size_t i = 0; i < a->size; i++) gsl_vector_complex_set(a, i, gsl_complex_mul(gsl_vector_complex_get(a, i), x)); return 0; }

/* real matrix, row-major */
typedef struct { size_t size1; size_t size2; size_t tda; double* data; gsl_block* block; int owner; } gsl_matrix;
static inline gsl_matrix* gsl_matrix_calloc(size_t n1, size_t n2) {
  gsl_matrix* m = (gsl_matrix*)malloc(sizeof(gsl_matrix));
  gsl_block* b = (gsl_block*)malloc(sizeof(gsl_block));
  b->size = n1 * n2; b->data = (double*)calloc(n1 * n2 ? n1 * n2 : 1, sizeof(double));
  m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = b->data; m->block = b; m->owner = 1; return m; }
static inline gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2) { return gsl_matrix_calloc(n1, n2); }
static inline void gsl_matrix_free(gsl_matrix* m) {
  if (!m) return; if (m->owner && m->block) { free(m->block->data); free(m->block); } free(m); }
static inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
static inline void gsl_matrix_set_zero(gsl_matrix* m) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = 0; }
static inline void gsl_matrix_set_all(gsl_matrix* m, double x) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = x; }
static inline void gsl_matrix_set_identity(gsl_matrix* m) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = (i == j) ? 1.0 : 0.0; }
static inline int gsl_matrix_memcpy(gsl_matrix* d, const gsl_matrix* s) {
  for (size_t i = 0; i < s->size1; i++) for (size_t j = 0; j < s->size2; j++) d->data[i * d->tda + j] = s->data[i * s->tda + j]; return 0; }
static inline int gsl_matrix_get_row(gsl_vector* v, const gsl_matrix* m, size_t i) {
  for (size_t j = 0; j < m->size2; j++) gsl_vector_set(v, j, gsl_matrix_get(m, i, j)); return 0; }
static inline int gsl_matrix_set_row(gsl_matrix* m, size_t i, const gsl_vector* v) {
  for (size_t j = 0; j < m->size2; j++) gsl_matrix_set(m, i, j, gsl_vector_get(v, j)); return 0; }
static inline int gsl_matrix_get_col(gsl_vector* v, const gsl_matrix* m, size_t j) {
  for (size_t i = 0; i < m->size1; i++) gsl_vector_set(v, i, gsl_matrix_get(m, i, j)); return 0; }
static inline int gsl_matrix_set_col(gsl_matrix* m, size_t j, const gsl_vector* v) {
  for (size_t i = 0; i < m->size1; i++) gsl_matrix_set(m, i, j, gsl_vector_get(v, i)); return 0; }
static inline int gsl_matrix_scale(gsl_matrix* m, double x) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] *= x; return 0; }
static inline int gsl_matrix_add(gsl_matrix* a, const gsl_matrix* b) {
  for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] += b->data[i * b->tda + j]; return 0; }

/* complex matrix, row-major, interleaved */
typedef struct { size_t size1; size_t size2; size_t tda; double* data; gsl_block_complex* block; int owner; } gsl_matrix_complex;
static inline gsl_matrix_complex* gsl_matrix_complex_calloc(size_t n1, size_t n2) {
  gsl_matrix_complex* m = (gsl_matrix_complex*)malloc(sizeof(gsl_matrix_complex));
  gsl_block_complex* b = (gsl_block_complex*)malloc(sizeof(gsl_block_complex));
  b->size = n1 * n2; b->data = (double*)calloc(2 * (n1 * n2 ? n1 * n2 : 1), sizeof(double));
  m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = b->data; m->block = b; m->owner = 1; return m; }
static inline gsl_matrix_complex* gsl_matrix_complex_alloc(size_t n1, size_t n2) { return gsl_matrix_complex_calloc(n1, n2); }
static inline void gsl_matrix_complex_free(gsl_matrix_complex* m) {
  if (!m) return; if (m->owner && m->block) { free(m->block->data); free(m->block); } free(m); }
static inline gsl_complex gsl_matrix_complex_get(const gsl_matrix_complex* m, size_t i, size_t j) {
  const double* p = m->data + 2 * (i * m->tda + j); return gsl_complex_rect(p[0], p[1]); }
static inline void gsl_matrix_complex_set(gsl_matrix_complex* m, size_t i, size_t j, gsl_complex z) {
  double* p = m->data + 2 * (i * m->tda + j); p[0] = z.dat[0]; p[1] = z.dat[1]; }
static inline void gsl_matrix_complex_set_zero(gsl_matrix_complex* m) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_complex_rect(0, 0)); }
static inline void gsl_matrix_complex_set_all(gsl_matrix_complex* m, gsl_complex z) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, z); }
static inline void gsl_matrix_complex_set_identity(gsl_matrix_complex* m) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_complex_rect(i == j ? 1.0 : 0.0, 0)); }
static inline int gsl_matrix_complex_memcpy(gsl_matrix_complex* d, const gsl_matrix_complex* s) {
  for (size_t i = 0; i < s->size1; i++) for (size_t j = 0; j < s->size2; j++) gsl_matrix_complex_set(d, i, j, gsl_matrix_complex_get(s, i, j)); return 0; }
static inline int gsl_matrix_complex_scale(gsl_matrix_complex* m, const gsl_complex x) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_complex_mul(gsl_matrix_complex_get(m, i, j), x)); return 0; }
static inline int gsl_matrix_complex_add(gsl_matrix_complex* a, const gsl_matrix_complex* b) {
  for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) gsl_matrix_complex_set(a, i, j, gsl_complex_add(gsl_matrix_complex_get(a, i, j), gsl_matrix_complex_get(b, i, j))); return 0; }
static inline int gsl_matrix_complex_sub(gsl_matrix_complex* a, const gsl_matrix_complex* b) {
  for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) gsl_matrix_complex_set(a, i, j, gsl_complex_sub(gsl_matrix_complex_get(a, i, j), gsl_matrix_complex_get(b, i, j))); return 0; }
static inline int gsl_matrix_complex_get_row(gsl_vector_complex* v, const gsl_matrix_complex* m, size_t i) {
  for (size_t j = 0; j < m->size2; j++) gsl_vector_complex_set(v, j, gsl_matrix_complex_get(m, i, j)); return 0; }
static inline int gsl_matrix_complex_set_row(gsl_matrix_complex* m, size_t i, const gsl_vector_complex* v) {
  for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_vector_complex_get(v, j)); return 0; }
static inline int gsl_matrix_complex_get_col(gsl_vector_complex* v, const gsl_matrix_complex* m, size_t j) {
  for (size_t i = 0; i < m->size1; i++) gsl_vector_complex_set(v, i, gsl_matrix_complex_get(m, i, j)); return 0; }
static inline int gsl_matrix_complex_set_col(gsl_matrix_complex* m, size_t j, const gsl_vector_complex* v) {
  for (size_t i = 0; i < m->size1; i++) gsl_matrix_complex_set(m, i, j, gsl_vector_complex_get(v, i)); return 0; }

/* ------------------------------------------------------------------ BLAS subset */
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef CBLAS_TRANSPOSE CBLAS_TRANSPOSE_t;

static inline gsl_complex shim_op_get_(CBLAS_TRANSPOSE tr, const gsl_matrix_complex* A, size_t i, size_t j) {
  /* element (i,j) of op(A) */
  if (tr == CblasNoTrans) return gsl_matrix_complex_get(A, i, j);
  if (tr == CblasTrans) return gsl_matrix_complex_get(A, j, i);
  return gsl_complex_conjugate(gsl_matrix_complex_get(A, j, i));
}
static inline int gsl_blas_zdotc(const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_complex* dotc) {
  double re = 0, im = 0;
  for (size_t i = 0; i < x->size; i++) {
    gsl_complex a = gsl_vector_complex_get(x, i), b = gsl_vector_complex_get(y, i);
    re += a.dat[0] * b.dat[0] + a.dat[1] * b.dat[1];
    im += a.dat[0] * b.dat[1] - a.dat[1] * b.dat[0];
  }
  dotc->dat[0] = re; dotc->dat[1] = im; return 0; }
static inline int gsl_blas_zdotu(const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_complex* dotu) {
  double re = 0, im = 0;
  for (size_t i = 0; i < x->size; i++) {
    gsl_complex a = gsl_vector_complex_get(x, i), b = gsl_vector_complex_get(y, i);
    re += a.dat[0] * b.dat[0] - a.dat[1] * b.dat[1];
    im += a.dat[0] * b.dat[1] + a.dat[1] * b.dat[0];
  }
  dotu->dat[0] = re; dotu->dat[1] = im; return 0; }
static inline double gsl_blas_dznrm2(const gsl_vector_complex* x) {
  double scale = 0.0, ssq = 1.0; /* scaled sum of squares */
  for (size_t i = 0; i < x->size; i++) for (int p = 0; p < 2; p++) {
    double a = fabs(x->data[2 * i * x->stride + p]);
    if (a != 0.0) { if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; } else ssq += (a / scale) * (a / scale); }
  }
  return scale * sqrt(ssq); }
static inline double gsl_blas_dnrm2(const gsl_vector* x) {
  double s = 0; for (size_t i = 0; i < x->size; i++) s += x->data[i * x->stride] * x->data[i * x->stride]; return sqrt(s); }
static inline int gsl_blas_ddot(const gsl_vector* x, const gsl_vector* y, double* r) {
  double s = 0; for (size_t i = 0; i < x->size; i++) s += x->data[i * x->stride] * y->data[i * y->stride]; *r = s; return 0; }
static inline int gsl_blas_zaxpy(const gsl_complex alpha, const gsl_vector_complex* x, gsl_vector_complex* y) {
  for (size_t i = 0; i < x->size; i++)
    gsl_vector_complex_set(y, i, gsl_complex_add(gsl_vector_complex_get(y, i), gsl_complex_mul(alpha, gsl_vector_complex_get(x, i))));
  return 0; }
static inline void gsl_blas_zscal(const gsl_complex alpha, gsl_vector_complex* x) {
  for (size_t i = 0; i < x->size; i++) gsl_vector_complex_set(x, i, gsl_complex_mul(alpha, gsl_vector_complex_get(x, i))); }
static inline void gsl_blas_zdscal(double alpha, gsl_vector_complex* x) {
  for (size_t i = 0; i < x->size; i++) gsl_vector_complex_set(x, i, gsl_complex_mul_real(gsl_vector_complex_get(x, i), alpha)); }
/* y = alpha op(A) x + beta y */
static inline int gsl_blas_zgemv(CBLAS_TRANSPOSE_t tr, const gsl_complex alpha, const gsl_matrix_complex* A,
                                 const gsl_vector_complex* x, const gsl_complex beta, gsl_vector_complex* y) {
  size_t rows = (tr == CblasNoTrans) ? A->size1 : A->size2;
  size_t cols = (tr == CblasNoTrans) ? A->size2 : A->size1;
  for (size_t i = 0; i < rows; i++) {
    gsl_complex acc = gsl_complex_rect(0, 0);
    for (size_t j = 0; j < cols; j++) acc = gsl_complex_add(acc, gsl_complex_mul(shim_op_get_(tr, A, i, j), gsl_vector_complex_get(x, j)));
    gsl_complex yi = gsl_complex_mul(beta, gsl_vector_complex_get(y, i));
    if (beta.dat[0] == 0.0 && beta.dat[1] == 0.0) yi = gsl_complex_rect(0, 0); /* BLAS: beta==0 ignores y */
    gsl_vector_complex_set(y, i, gsl_complex_add(yi, gsl_complex_mul(alpha, acc)));
  }
  return 0; }
/* A += alpha x y^T */
static inline int gsl_blas_zgeru(const gsl_complex alpha, const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_matrix_complex* A) {
  for (size_t i = 0; i < A->size1; i++) for (size_t j = 0; j < A->size2; j++)
    gsl_matrix_complex_set(A, i, j, gsl_complex_add(gsl_matrix_complex_get(A, i, j),
      gsl_complex_mul(alpha, gsl_complex_mul(gsl_vector_complex_get(x, i), gsl_vector_complex_get(y, j)))));
  return 0; }
/* A += alpha x y^H */
static inline int gsl_blas_zgerc(const gsl_complex alpha, const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_matrix_complex* A) {
  for (size_t i = 0; i < A->size1; i++) for (size_t j = 0; j < A->size2; j++)
    gsl_matrix_complex_set(A, i, j, gsl_complex_add(gsl_matrix_complex_get(A, i, j),
      gsl_complex_mul(alpha, gsl_complex_mul(gsl_vector_complex_get(x, i), gsl_complex_conjugate(gsl_vector_complex_get(y, j))))));
  return 0; }
/* C = alpha op(A) op(B) + beta C */
static inline int gsl_blas_zgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, const gsl_complex alpha, const gsl_matrix_complex* A,
                                 const gsl_matrix_complex* B, const gsl_complex beta, gsl_matrix_complex* C) {
  size_t n1 = C->size1, n2 = C->size2, kk = (ta == CblasNoTrans) ? A->size2 : A->size1;
  gsl_matrix_complex* T = gsl_matrix_complex_calloc(n1, n2);
  for (size_t i = 0; i < n1; i++) for (size_t j = 0; j < n2; j++) {
    gsl_complex acc = gsl_complex_rect(0, 0);
    for (size_t k = 0; k < kk; k++) acc = gsl_complex_add(acc, gsl_complex_mul(shim_op_get_(ta, A, i, k), shim_op_get_(tb, B, k, j)));
    gsl_complex cij = gsl_complex_mul(beta, gsl_matrix_complex_get(C, i, j));
    if (beta.dat[0] == 0.0 && beta.dat[1] == 0.0) cij = gsl_complex_rect(0, 0);
    gsl_matrix_complex_set(T, i, j, gsl_complex_add(cij, gsl_complex_mul(alpha, acc)));
  }
  gsl_matrix_complex_memcpy(C, T); gsl_matrix_complex_free(T); return 0; }

/* ------------------------------------------------------------------ radix-2 complex DFT, packed (re,im) array */
typedef double* gsl_complex_packed_array;
static inline int shim_fft_radix2_(double* data, size_t stride, size_t n, int sign) {
  if (n == 0 || (n & (n - 1))) { fprintf(stderr, "gsl_shim: radix2 FFT length %zu is not a power of 2\n", n); return 1; }
  /* bit reversal */
  for (size_t i = 0, j = 0; i < n; i++) {
    if (i < j) {
      double tr = data[2 * stride * i], ti = data[2 * stride * i + 1];
      data[2 * stride * i] = data[2 * stride * j]; data[2 * stride * i + 1] = data[2 * stride * j + 1];
      data[2 * stride * j] = tr; data[2 * stride * j + 1] = ti;
    }
    size_t bit = n >> 1;
    while (bit && (j & bit)) { j ^= bit; bit >>= 1; }
    j |= bit;
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t half = len >> 1;
    for (size_t k = 0; k < half; k++) {
      double ang = sign * 2.0 * M_PI * (double)k / (double)len;
      double wr = cos(ang), wi = sin(ang);
      for (size_t s = k; s < n; s += len) {
        size_t a = 2 * stride * s, b = 2 * stride * (s + half);
        double xr = data[b] * wr - data[b + 1] * wi, xi = data[b] * wi + data[b + 1] * wr;
        data[b] = data[a] - xr; data[b + 1] = data[a + 1] - xi;
        data[a] += xr; data[a + 1] += xi;
      }
    }
  }
  return 0;
}
static inline int gsl_fft_complex_radix2_forward(gsl_complex_packed_array d, size_t stride, size_t n) { return shim_fft_radix2_(d, stride, n, -1); }
static inline int gsl_fft_complex_radix2_backward(gsl_complex_packed_array d, size_t stride, size_t n) { return shim_fft_radix2_(d, stride, n, +1); }
static inline int gsl_fft_complex_radix2_inverse(gsl_complex_packed_array d, size_t stride, size_t n) {
  int r = shim_fft_radix2_(d, stride, n, +1);
  for (size_t i = 0; i < n; i++) { d[2 * stride * i] /= (double)n; d[2 * stride * i + 1] /= (double)n; }
  return r; }

/* ------------------------------------------------------------------ special functions */
static inline double gsl_sf_sinc(double x) { double y = M_PI * x; return (fabs(y) < 1e-12) ? 1.0 : sin(y) / y; }

/* ------------------------------------------------------------------ complex Cholesky (A = L L^H, lower), for the WPE "next" row */
static inline int gsl_linalg_complex_cholesky_decomp(gsl_matrix_complex* A) {
  size_t n = A->size1;
  for (size_t j = 0; j < n; j++) {
    double d = GSL_REAL(gsl_matrix_complex_get(A, j, j));
    for (size_t k = 0; k < j; k++) d -= gsl_complex_abs2(gsl_matrix_complex_get(A, j, k));
    if (d <= 0.0) { fprintf(stderr, "gsl_shim: cholesky: matrix not positive definite\n"); return 1; }
    d = sqrt(d);
    gsl_matrix_complex_set(A, j, j, gsl_complex_rect(d, 0));
    for (size_t i = j + 1; i < n; i++) {
      gsl_complex s = gsl_matrix_complex_get(A, i, j);
      for (size_t k = 0; k < j; k++) s = gsl_complex_sub(s, gsl_complex_mul(gsl_matrix_complex_get(A, i, k), gsl_complex_conjugate(gsl_matrix_complex_get(A, j, k))));
      s = gsl_complex_div_real(s, d);
      gsl_matrix_complex_set(A, i, j, s);
      gsl_matrix_complex_set(A, j, i, gsl_complex_conjugate(s)); /* GSL stores L^H in the upper triangle */
    }
  }
  return 0; }
static inline int gsl_linalg_complex_cholesky_svx(const gsl_matrix_complex* LLT, gsl_vector_complex* x) {
  size_t n = LLT->size1;
  for (size_t i = 0; i < n; i++) { /* L y = b */
    gsl_complex s = gsl_vector_complex_get(x, i);
    for (size_t k = 0; k < i; k++) s = gsl_complex_sub(s, gsl_complex_mul(gsl_matrix_complex_get(LLT, i, k), gsl_vector_complex_get(x, k)));
    gsl_vector_complex_set(x, i, gsl_complex_div_real(s, GSL_REAL(gsl_matrix_complex_get(LLT, i, i))));
  }
  for (size_t ii = n; ii-- > 0;) { /* L^H x = y */
    gsl_complex s = gsl_vector_complex_get(x, ii);
    for (size_t k = ii + 1; k < n; k++) s = gsl_complex_sub(s, gsl_complex_mul(gsl_complex_conjugate(gsl_matrix_complex_get(LLT, k, ii)), gsl_vector_complex_get(x, k)));
    gsl_vector_complex_set(x, ii, gsl_complex_div_real(s, GSL_REAL(gsl_matrix_complex_get(LLT, ii, ii))));
  }
  return 0; }
static inline int gsl_linalg_complex_cholesky_solve(const gsl_matrix_complex* LLT, const gsl_vector_complex* b, gsl_vector_complex* x) {
  gsl_vector_complex_memcpy(x, b); return gsl_linalg_complex_cholesky_svx(LLT, x); }

#endif /* ORACLE_GSL_SHIM_H */
