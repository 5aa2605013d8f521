// btkb_fft.cuh — register-resident Stockham FFT used by the OverSampledDFT analysis / synthesis kernels (sm_100a).
//
// An M-point complex DFT is done by NT = M/8 threads holding 8 complex values each.  M = R0 * 8^P with R0 in {2,4,8}:
// one leading radix-R0 pass (no twiddles) followed by P radix-8 passes.  Between passes the data goes through a
// ping-pong shared-memory buffer (one barrier per pass).  The per-thread twiddles depend only on the thread index, so
// they are computed once per CTA (sincospif, exact argument reduction) and stay in registers for every transform the
// CTA performs.
//
// Direction: SIGN = +1 is the reference's gsl_fft_complex_radix2_backward (unnormalised e^{+2 pi i nk/M}, used by
// OverSampledDFTAnalysisBank::next, modulated.cc:396); SIGN = -1 is gsl_fft_complex_radix2_forward (synthesis,
// modulated.cc:559).
//
// PK = true selects the packed 2 x fp32 forms of btkb_f2.cuh (FADD2 / FMUL2 / FFMA2): same operations and roundings per component,
// half the issue slots for the complex arithmetic (a radix-8 butterfly is 28 instructions instead of 56, a twiddle product 2
// instead of 4).  Bit-identical to PK = false (tests/test_fft_packed_host.py runs both on the CPU).
#pragma once
#include <cuda_runtime.h>
#include "btkb_f2.cuh"

namespace btkb {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ float2 mul_si(float2 a) {
  return SIGN > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

#if defined(__CUDACC__)   // (this header is also compiled for the host by tests/host/fft_packed_host.cc)
// Barrier over the NT threads of ONE transform group (group index grp of G per CTA): the groups of a CTA work on different frames with
// private exchange buffers, so inside a transform they only have to wait for their own threads — a warp-level barrier when a group is
// one warp, a named barrier (ids 1..G; 0 stays __syncthreads) when it is several, the CTA barrier when the CTA is one group.
template <int NT, int G>
__device__ __forceinline__ void group_sync(int grp) {
  if constexpr (G == 1) __syncthreads();
  else if constexpr (NT <= 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(NT) : "memory");
}
#endif

template <int M>
struct FftPlan {
  static_assert(M >= 64 && M <= 4096 && (M & (M - 1)) == 0, "M must be a power of two in [64,4096]");
  static constexpr int log2m() { int q = 0; for (int v = M; v > 1; v >>= 1) q++; return q; }
  static constexpr int P = (log2m() - 1) / 3;                           // number of radix-8 passes
  __host__ __device__ static constexpr int pow8(int p) { return p == 0 ? 1 : 8 * pow8(p - 1); }
  static constexpr int R0 = M / pow8(P);                                // leading radix: 2, 4 or 8
  static constexpr int NT = M / 8;                                      // threads per transform
  static constexpr int NPASS = P + 1;
  static constexpr int BUF = M + M / 16;                                // padded buffer length (float2), covers every layout of lidx<>
};

// padded shared-memory index: breaks the stride-8 / stride-64 patterns of the pass outputs
__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }

// Per-exchange shared-memory layouts (a store and the load that consumes it must agree; different exchanges may differ):
//   0 identity            for stores with stride >= 16 (lanes consecutive) and the natural-order spectrum
//   1 i + (i >> 4)        for the leading pass (lane stride R0 <= 8 elements)
//   2 i ^ (bit6(i) << 3)  for the radix-8 pass with Ns = 8 (lanes (a, b) -> 64 a + b): the two a-groups of a half warp
//                         land 8 slots apart, which the i >> 4 padding does not achieve (2-way conflicts, ncu r01b)
template <int LAY>
__device__ __forceinline__ int lidx(int i) { return LAY == 0 ? i : (LAY == 1 ? i + (i >> 4) : (i ^ (((i >> 6) & 1) << 3))); }
__host__ __device__ constexpr int lay_of(int Ns) { return Ns >= 16 ? 0 : (Ns == 8 ? 2 : 1); }

// The same indices in "per-thread base + compile-time offset" form (used by the PK = true code: the eight addresses of an exchange
// then cost one or two registers and no per-iteration integer arithmetic).
//   lidx_ld<LAY, S>(tg, r)  == lidx<LAY>(tg + r S)               loads of a radix-8 pass, S = M/8, tg < S
//   lidx_st<LAY, Ns>(tg, r) == lidx<LAY>(j0 + r Ns), j0 = (tg - tg % Ns) 8 + tg % Ns     Stockham scatter of a radix-8 pass
template <int LAY, int S>
__device__ __forceinline__ int lidx_ld(int tg, int r) {
  if constexpr (LAY == 0) return tg + r * S;
  else if constexpr (LAY == 1) {
    static_assert(S % 16 == 0, "stride must be a multiple of the padding period");
    return (tg + (tg >> 4)) + r * (S + S / 16);            // (tg + r S) >> 4 == (tg >> 4) + r S / 16
  } else {
    static_assert(S == 64, "the XOR layout is only used for M = 512");
    return ((r & 1) ? (tg ^ 8) : tg) + r * S;               // bit 6 of tg + 64 r is r & 1 (tg < 64); the XOR touches bit 3 only
  }
}
template <int LAY, int Ns>
__device__ __forceinline__ int lidx_st(int tg, int r) {
  const int k = tg % Ns;
  const int j0 = (tg - k) * 8 + k;
  if constexpr (LAY == 0) return j0 + r * Ns;
  else if constexpr (LAY == 1) {
    static_assert(Ns == 2 || Ns == 4, "padded layout: strides below 8");
    return (j0 + (j0 >> 4)) + r * Ns + ((r * Ns) >> 4);     // j0 - k is a multiple of 16 and k + r Ns crosses 16 exactly when r Ns does
  } else {
    static_assert(Ns == 8, "XOR layout: stride 8");
    const int s8 = ((tg >> 3) & 1) << 3;                   // bit 6 of j0 + 8 r is bit 3 of tg; bit 3 of j0 + 8 r is r & 1
    return ((r & 1) ? (j0 - s8) : (j0 + s8)) + r * Ns;
  }
}

template <int SIGN, bool PK = false>
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  float2 t = a;
  if constexpr (PK) { a = f2_add(t, b); b = f2_sub(t, b); }
  else { a = cadd(t, b); b = csub(t, b); }
}
template <int SIGN, bool PK = false>
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  if constexpr (PK) {
    // a3 = SIGN i (v1 - v3) is never formed: v1 = a1 + SIGN i d, v3 = a1 - SIGN i d (swap + half negation are operand modifiers)
    const float2 a0 = f2_add(v0, v2), a1 = f2_sub(v0, v2), a2 = f2_add(v1, v3), d = f2_sub(v1, v3);
    v0 = f2_add(a0, a2); v2 = f2_sub(a0, a2); v1 = f2_add_ib<SIGN>(a1, d); v3 = f2_sub_ib<SIGN>(a1, d);
    return;
  }
  float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_si<SIGN>(csub(v1, v3));
  v0 = cadd(a0, a2); v2 = csub(a0, a2); v1 = cadd(a1, a3); v3 = csub(a1, a3);
}
// X[q] = sum_r v[r] exp(SIGN 2 pi i q r / 8), natural order in and out
template <int SIGN, bool PK = false>
__device__ __forceinline__ void dft8(float2* v) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4<SIGN, PK>(e0, e1, e2, e3);
  dft4<SIGN, PK>(o0, o1, o2, o3);
  const float s = 0.70710678118654752440f;
  // o1 (1 + SIGN i)/sqrt2 and o3 (-1 + SIGN i)/sqrt2 are never formed: the scale by 1/sqrt2 is folded into the last butterfly as an
  // explicit multiply-add, v[1] = fma(o1 + SIGN i o1, s, e1), v[5] = fma(o1 + SIGN i o1, -s, e1) (likewise v[3], v[7]).  nvcc's
  // default -fmad=true contracted the scalar "(sum) * s" + "e +- that" into exactly these FFMAs already (8 per butterfly); writing
  // them out makes the packed form (FFMA2) round identically and makes the host build of this header (no contraction) agree
  // with the device.
  if constexpr (PK) {
    const float2 u1 = f2_add_ib<SIGN>(o1, o1);
    const float2 u3 = f2_add_ib<SIGN>(make_float2(-o3.x, -o3.y), o3);
    v[0] = f2_add(e0, o0); v[4] = f2_sub(e0, o0);
    v[1] = f2_fma_s(u1, s, e1); v[5] = f2_fma_s(u1, -s, e1);
    v[2] = f2_add_ib<SIGN>(e2, o2); v[6] = f2_sub_ib<SIGN>(e2, o2);
    v[3] = f2_fma_s(u3, s, e3); v[7] = f2_fma_s(u3, -s, e3);
    return;
  }
  // o2 *= SIGN i
  const float2 t1 = mul_si<SIGN>(o1);
  const float2 u1 = make_float2(o1.x + t1.x, o1.y + t1.y);
  o2 = mul_si<SIGN>(o2);
  const float2 t3 = mul_si<SIGN>(o3);
  const float2 u3 = make_float2(t3.x - o3.x, t3.y - o3.y);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = make_float2(fmaf(u1.x, s, e1.x), fmaf(u1.y, s, e1.y)); v[5] = make_float2(fmaf(u1.x, -s, e1.x), fmaf(u1.y, -s, e1.y));
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = make_float2(fmaf(u3.x, s, e3.x), fmaf(u3.y, s, e3.y)); v[7] = make_float2(fmaf(u3.x, -s, e3.x), fmaf(u3.y, -s, e3.y));
}

// Per-thread twiddles for the P radix-8 passes: tw[p][r-1] = exp(SIGN 2 pi i k r / (Ns 8)), k = tid % Ns, Ns = R0 8^p.
template <int M, int SIGN>
struct FftTwiddles {
  using Plan = FftPlan<M>;
  float2 tw[Plan::P][7];
  __device__ __forceinline__ void init(int tid) {
    int Ns = Plan::R0;
#pragma unroll
    for (int p = 0; p < Plan::P; p++) {
      int k = tid % Ns;
#pragma unroll
      for (int r = 1; r < 8; r++) {
        float s, c;
        // angle / pi = SIGN * 2 k r / (8 Ns)  (exact in float: k r < 2^15, power-of-two denominator)
        sincospif((float)(SIGN * 2 * k * r) / (float)(8 * Ns), &s, &c);
        tw[p][r - 1] = make_float2(c, s);
      }
      Ns *= 8;
    }
  }
  // Same values from a table tab[n] = exp(+2 pi i n / M), n < M, built once on the host in double precision: 7 P loads
  // instead of 7 P sincospif evaluations (~40 instructions each) in every CTA prologue.
  __device__ __forceinline__ void init_from_table(int tid, const float2* __restrict__ tab) {
    int Ns = Plan::R0;
#pragma unroll
    for (int p = 0; p < Plan::P; p++) {
      const int k = tid % Ns;
      const int step = M / (8 * Ns);
#pragma unroll
      for (int r = 1; r < 8; r++) {
        float2 w = __ldg(tab + k * r * step);     // k r < 8 Ns  =>  index < M
        tw[p][r - 1] = (SIGN > 0) ? w : make_float2(w.x, -w.y);
      }
      Ns *= 8;
    }
  }
};

// Store the leading-pass or radix-8-pass results (Stockham scatter): element r of butterfly j goes to
// (j / Ns) Ns R + (j % Ns) + r Ns.
template <int R>
__device__ __forceinline__ void stockham_store(float2* buf, const float2* v, int j, int Ns) {
  int k = j % Ns;
  int j0 = (j - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; r++) buf[pidx(j0 + r * Ns)] = v[r];
}

// The leading radix-R0 pass on registers that were loaded as v[b*R0 + r] = in[(tid + b NT) + r M/R0].
template <int M, int SIGN, bool PK = false>
__device__ __forceinline__ void fft_first_pass(float2* v, float2* buf, int tid) {
  using Plan = FftPlan<M>;
  constexpr int R0 = Plan::R0, NT = Plan::NT, NB = 8 / R0;
#pragma unroll
  for (int b = 0; b < NB; b++) {
    float2* w = v + b * R0;
    if (R0 == 8) dft8<SIGN, PK>(w);
    else if (R0 == 4) dft4<SIGN, PK>(w[0], w[1], w[2], w[3]);
    else dft2<SIGN, PK>(w[0], w[1]);
    if constexpr (PK) {   // pidx(j R0 + r) == j R0 + j / (16 / R0) + r  (R0 (j mod 16/R0) + r <= 15)
      const int j = tid + b * NT;
      float2* dst = buf + (j * R0 + j / (16 / R0));
#pragma unroll
      for (int r = 0; r < R0; r++) dst[r] = w[r];
    } else
    stockham_store<R0>(buf, w, tid + b * NT, 1);
  }
}

// Radix-8 pass p (0-based among the radix-8 passes): reads `in`, writes `out` (both padded smem buffers).
template <int M, int SIGN>
__device__ __forceinline__ void fft_radix8_pass(const float2* in, float2* out, int tid, int p, int Ns,
                                                const FftTwiddles<M, SIGN>& T) {
  float2 v[8];
#pragma unroll
  for (int r = 0; r < 8; r++) v[r] = in[pidx(tid + r * (M / 8))];
#pragma unroll
  for (int r = 1; r < 8; r++) v[r] = cmul(v[r], T.tw[p][r - 1]);
  dft8<SIGN>(v);
  stockham_store<8>(out, v, tid, Ns);
}

// Full transform after the caller filled v (leading-pass operand order).  bufA/bufB: two padded smem buffers of
// FftPlan<M>::BUF float2 each, private to the NT threads of this transform.  `sync()` must be a barrier covering at
// least those NT threads.  On return the natural-order spectrum is in the returned buffer (already synchronised).
template <int M, int SIGN, typename Sync>
__device__ __forceinline__ float2* fft_run(float2* v, float2* bufA, float2* bufB, int tid,
                                           const FftTwiddles<M, SIGN>& T, Sync sync) {
  using Plan = FftPlan<M>;
  fft_first_pass<M, SIGN>(v, bufA, tid);
  sync();
  float2* in = bufA;
  float2* out = bufB;
  int Ns = Plan::R0;
#pragma unroll
  for (int p = 0; p < Plan::P; p++) {
    fft_radix8_pass<M, SIGN>(in, out, tid, p, Ns, T);
    sync();
    float2* t = in; in = out; out = t;
    Ns *= 8;
  }
  return in;
}


// One radix-8 pass (index PASS among the P radix-8 passes) of TWO independent transforms held in v0 / v1, exchanged through
// their own in-place buffers b0 / b1.  The buffers were filled by the previous exchange with layout lay_of(previous stride).
// On the last pass only the upper half of the spectrum (indices >= M/2, i.e. r = 4..7) is written back, in natural order:
// callers that untangle real sequences keep Z[k] for k < M/2 in registers (v[r], r < 4) and fetch only Z[M - k].
template <int M, int SIGN, int PASS, typename Sync, bool PK = false>
__device__ __forceinline__ void fft_pass_pair(float2* v0, float2* v1, float2* b0, float2* b1, int tg, const FftTwiddles<M, SIGN>& T, Sync sync) {
  using Plan = FftPlan<M>;
  constexpr int NsPrev = (PASS == 0) ? 1 : Plan::R0 * Plan::pow8(PASS - 1);
  constexpr int Ns = Plan::R0 * Plan::pow8(PASS);
  constexpr int LIN = (PASS == 0) ? 1 : lay_of(NsPrev);
  constexpr int LOUT = lay_of(Ns);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if constexpr (PK) { v0[r] = b0[lidx_ld<LIN, M / 8>(tg, r)]; v1[r] = b1[lidx_ld<LIN, M / 8>(tg, r)]; }
    else { v0[r] = b0[lidx<LIN>(tg + r * (M / 8))]; v1[r] = b1[lidx<LIN>(tg + r * (M / 8))]; }
  }
  sync();
#pragma unroll
  for (int r = 1; r < 8; r++) {
    if constexpr (PK) { v0[r] = f2_cmul(v0[r], T.tw[PASS][r - 1]); v1[r] = f2_cmul(v1[r], T.tw[PASS][r - 1]); }
    else { v0[r] = cmul(v0[r], T.tw[PASS][r - 1]); v1[r] = cmul(v1[r], T.tw[PASS][r - 1]); }
  }
  dft8<SIGN, PK>(v0);
  dft8<SIGN, PK>(v1);
  if (PASS == Plan::P - 1) {
    static_assert(Ns * 8 == M || PASS != Plan::P - 1, "last pass stride");
#pragma unroll
    for (int r = 4; r < 8; r++) { b0[tg + r * (M / 8)] = v0[r]; b1[tg + r * (M / 8)] = v1[r]; }
  } else if constexpr (PK) {
#pragma unroll
    for (int r = 0; r < 8; r++) { b0[lidx_st<LOUT, Ns>(tg, r)] = v0[r]; b1[lidx_st<LOUT, Ns>(tg, r)] = v1[r]; }
  } else {
    const int k = tg % Ns;
    const int j0 = (tg - k) * 8 + k;
#pragma unroll
    for (int r = 0; r < 8; r++) { b0[lidx<LOUT>(j0 + r * Ns)] = v0[r]; b1[lidx<LOUT>(j0 + r * Ns)] = v1[r]; }
  }
  sync();
}
template <int M, int SIGN, int PASS, typename Sync, bool PK = false>
struct FftPassChain {
  static __device__ __forceinline__ void run(float2* v0, float2* v1, float2* b0, float2* b1, int tg, const FftTwiddles<M, SIGN>& T, Sync sync) {
    if constexpr (PASS < FftPlan<M>::P) {
      fft_pass_pair<M, SIGN, PASS, Sync, PK>(v0, v1, b0, b1, tg, T, sync);
      FftPassChain<M, SIGN, PASS + 1, Sync, PK>::run(v0, v1, b0, b1, tg, T, sync);
    }
  }
};

}  // namespace btkb
