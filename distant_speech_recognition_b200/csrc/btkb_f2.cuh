// btkb_f2.cuh — packed 2 x fp32 arithmetic on float2 values (sm_100a: FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2).
//
// The filter-bank kernels are bound by instruction issue, not by the FMA pipe (ncu: issue slots 64 % busy, FMA pipe 38 %), and
// more than half of what they issue is fp32 arithmetic on complex values, i.e. on register PAIRS.  Blackwell's packed fp32
// instructions do both halves of a pair in one issue slot, and ptxas folds the operand patterns complex arithmetic needs into
// operand modifiers (checked with cuobjdump on this toolchain, nvcc 12.9):
//     R.F32            one register broadcast to both halves          (scaling a complex value by a real tap / twiddle part)
//     R.F32x2.LO_HI    the pair with its halves swapped               (multiplication by +-i)
//     [-]R.F32x2...NP  negation of one half only                      (conjugation, the cross terms of a complex product; accepted on
//                      the first multiplicand and on the addend, NOT on the second multiplicand — see f2_cmac)
// so a complex add is 1 instruction instead of 2, a complex multiply 2 instead of 4, "a + i b" 1 instead of 2 — with NO extra
// registers and no data movement: the float2 values are already even/odd register pairs.
//
// Every function below performs, per component, exactly the IEEE operations (same order, same roundings) of the scalar code it
// replaces in btkb_fft.cuh; negations and swaps are exact.  Results are therefore bit-identical to the scalar path, which is what
// tests/test_fft_packed_host.py checks on the CPU: compiled for the host (no __CUDA_ARCH__), the same functions are plain scalar
// code with fmaf(), and the packed and scalar FFTs must agree bit for bit.
#pragma once
#include <cuda_runtime.h>
#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace btkb {

#if defined(__CUDACC__)
#define BTKB_F2 __host__ __device__ __forceinline__
#else
#define BTKB_F2 inline
#endif

#if defined(__CUDA_ARCH__)
typedef unsigned long long f2raw;
__device__ __forceinline__ f2raw f2_pk(float lo, float hi) { f2raw r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 f2_upk(f2raw v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f2raw f2_add_raw(f2raw a, f2raw b) { f2raw d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2raw f2_mul_raw(f2raw a, f2raw b) { f2raw d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2raw f2_fma_raw(f2raw a, f2raw b, f2raw c) { f2raw d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#endif

// (a.x + b.x, a.y + b.y)
BTKB_F2 float2 f2_add(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_add_raw(f2_pk(a.x, a.y), f2_pk(b.x, b.y)));
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
// (a.x - b.x, a.y - b.y)
BTKB_F2 float2 f2_sub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_add_raw(f2_pk(a.x, a.y), f2_pk(-b.x, -b.y)));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a + SIGN i b  =  SIGN > 0 ? (a.x - b.y, a.y + b.x) : (a.x + b.y, a.y - b.x)
template <int SIGN>
BTKB_F2 float2 f2_add_ib(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return SIGN > 0 ? f2_upk(f2_add_raw(f2_pk(a.x, a.y), f2_pk(-b.y, b.x))) : f2_upk(f2_add_raw(f2_pk(a.x, a.y), f2_pk(b.y, -b.x)));
#else
  return SIGN > 0 ? make_float2(a.x + (-b.y), a.y + b.x) : make_float2(a.x + b.y, a.y + (-b.x));
#endif
}
// a - SIGN i b
template <int SIGN>
BTKB_F2 float2 f2_sub_ib(float2 a, float2 b) { return f2_add_ib<-SIGN>(a, b); }
// (a.x s, a.y s)
BTKB_F2 float2 f2_scale(float2 a, float s) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_mul_raw(f2_pk(a.x, a.y), f2_pk(s, s)));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}
// (fma(a.x, s, c.x), fma(a.y, s, c.y)) — a real tap times a complex sample, accumulated
BTKB_F2 float2 f2_fma_s(float2 a, float s, float2 c) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_fma_raw(f2_pk(a.x, a.y), f2_pk(s, s), f2_pk(c.x, c.y)));
#else
  return make_float2(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y));
#endif
}
// (fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)) — two independent MAC chains side by side
BTKB_F2 float2 f2_fma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_fma_raw(f2_pk(a.x, a.y), f2_pk(b.x, b.y), f2_pk(c.x, c.y)));
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// a + conj(b) = (a.x + b.x, a.y - b.y)   and   a - conj(b) = (a.x - b.x, a.y + b.y): the two sums that untangle a pair of real
// sequences from one complex transform
BTKB_F2 float2 f2_add_conj(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_add_raw(f2_pk(a.x, a.y), f2_pk(b.x, -b.y)));
#else
  return make_float2(a.x + b.x, a.y + (-b.y));
#endif
}
BTKB_F2 float2 f2_sub_conj(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_add_raw(f2_pk(a.x, a.y), f2_pk(-b.x, b.y)));
#else
  return make_float2(a.x + (-b.x), a.y + b.y);
#endif
}
// -i d s = (d.y s, -d.x s)
BTKB_F2 float2 f2_scale_mi(float2 d, float s) {
#if defined(__CUDA_ARCH__)
  return f2_upk(f2_mul_raw(f2_pk(d.y, -d.x), f2_pk(s, s)));
#else
  return make_float2(d.y * s, (-d.x) * s);
#endif
}
// ---- complex multiply-accumulates of the per-bin kernels (btkb_perbin.cu cmac / cmac_conj and the NLMS update), two FFMA2 each.
// The host forms spell out what each half of the two packed instructions computes.
// acc + a b:        x: fma(a.y, -b.y, fma(a.x, b.x, acc.x))     y: fma(a.y, b.x, fma(a.x, b.y, acc.y))
BTKB_F2 float2 f2_cmac(float2 acc, float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  // (the swapped / half-negated pair goes FIRST: ptxas folds those patterns into modifiers of the first multiplicand only, a
  // pair in the second slot is rebuilt with a MOV and an FADD per use; the broadcast is accepted in either slot)
  const f2raw t = f2_fma_raw(f2_pk(b.x, b.y), f2_pk(a.x, a.x), f2_pk(acc.x, acc.y));
  return f2_upk(f2_fma_raw(f2_pk(-b.y, b.x), f2_pk(a.y, a.y), t));
#else
  return make_float2(fmaf(a.y, -b.y, fmaf(a.x, b.x, acc.x)), fmaf(a.y, b.x, fmaf(a.x, b.y, acc.y)));
#endif
}
// acc + a conj(b):  x: fma(a.y, b.y, fma(a.x, b.x, acc.x))      y: fma(-a.x, b.y, fma(a.y, b.x, acc.y))
BTKB_F2 float2 f2_cmac_conj(float2 acc, float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  const f2raw t = f2_fma_raw(f2_pk(a.x, a.y), f2_pk(b.x, b.x), f2_pk(acc.x, acc.y));
  return f2_upk(f2_fma_raw(f2_pk(a.y, -a.x), f2_pk(b.y, b.y), t));
#else
  return make_float2(fmaf(a.y, b.y, fmaf(a.x, b.x, acc.x)), fmaf(-a.x, b.y, fmaf(a.y, b.x, acc.y)));
#endif
}
// x - c v  (the projector step of the NLMS update, c = C Yc):
//                   x: fma(-c.x, v.x, fma(c.y, v.y, x.x))       y: fma(-c.x, v.y, fma(c.y, -v.x, x.y))
BTKB_F2 float2 f2_sub_cmul(float2 x, float2 c, float2 v) {
#if defined(__CUDA_ARCH__)
  const f2raw t = f2_fma_raw(f2_pk(v.y, -v.x), f2_pk(c.y, c.y), f2_pk(x.x, x.y));
  return f2_upk(f2_fma_raw(f2_pk(v.x, v.y), f2_pk(-c.x, -c.x), t));
#else
  return make_float2(fmaf(-c.x, v.x, fmaf(c.y, v.y, x.x)), fmaf(-c.x, v.y, fmaf(c.y, -v.x, x.y)));
#endif
}
// k + e conj(q)  (k = keep u):
//                   x: fma(e.y, q.y, fma(e.x, q.x, k.x))        y: fma(-e.x, q.y, fma(e.y, q.x, k.y))
BTKB_F2 float2 f2_add_cmulc(float2 k, float2 e, float2 q) {
#if defined(__CUDA_ARCH__)
  const f2raw t = f2_fma_raw(f2_pk(e.x, e.y), f2_pk(q.x, q.x), f2_pk(k.x, k.y));
  return f2_upk(f2_fma_raw(f2_pk(e.y, -e.x), f2_pk(q.y, q.y), t));
#else
  return make_float2(fmaf(e.y, q.y, fmaf(e.x, q.x, k.x)), fmaf(-e.x, q.y, fmaf(e.y, q.x, k.y)));
#endif
}
// a conj(b), the roundings of btkb::cmulc:  x: fma(a.x, b.x, a.y b.y)     y: fma(a.y, b.x, (-a.x) b.y)
BTKB_F2 float2 f2_cmulc(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  const f2raw t = f2_mul_raw(f2_pk(a.y, -a.x), f2_pk(b.y, b.y));
  return f2_upk(f2_fma_raw(f2_pk(a.x, a.y), f2_pk(b.x, b.x), t));
#else
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, (-a.x) * b.y));
#endif
}
// complex product, the roundings of btkb::cmul: (fma(a.x, w.x, -(a.y w.y)), fma(a.x, w.y, a.y w.x))
BTKB_F2 float2 f2_cmul(float2 a, float2 w) {
#if defined(__CUDA_ARCH__)
  const float2 t = f2_upk(f2_mul_raw(f2_pk(a.y, a.y), f2_pk(w.y, w.x)));
  return f2_upk(f2_fma_raw(f2_pk(a.x, a.x), f2_pk(w.x, w.y), f2_pk(-t.x, t.y)));
#else
  return make_float2(fmaf(a.x, w.x, -(a.y * w.y)), fmaf(a.x, w.y, a.y * w.x));
#endif
}

}  // namespace btkb
