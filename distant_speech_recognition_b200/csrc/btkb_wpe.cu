// btkb_wpe.cu — multi-channel WPE dereverberation between the analysis bank and the beamformer (sm_100a).
//
// Replaces MultiChannelWPEDereverberation::{fill_buffer_, calc_Thetan_, calc_Rr_, load_R_, estimate_Gn_,
// calc_every_channel_output} (dereverberation/dereverberation.cc:441-700).  Unit of work: a "problem" = one (utterance, bin)
// series of C channels x T frames; problems are independent, so the batch is a grid of them.
//
//   k_wpe_gather   X[T][C][Gp] -> S[g][c][Ts]      series-major copy (every later kernel reads contiguous series)
//   k_wpe_resid<0> theta[g][c][t] = max(|x_c(t) - G_c^H lags(t)|, 1e-3)^2                       calc_Thetan_  (:619-646)
//   k_wpe_corr     R_c = sum_s lags lags^H / theta_c(s) (lower triangle) and r_c^H as an extra row  calc_Rr_    (:553-617)
//   k_wpe_chol     + diagonal_bias, load_R_ (:648-663), Cholesky, solve R_c g = r_c              estimate_Gn_ (:665-690)
//   k_wpe_resid<1> X'[t][c][g] = x_c(t) - G_c^H lags(t)                                          calc_every_channel_output (:441-497)
//
// lags(s)[(c', l)] = x_c'(s - lower - l), channel-major then lag (get_lags_, :536-551).  The lag matrix is block-Toeplitz in
// the series, so k_wpe_corr keeps only the C series (and the C weight series 1/theta_c) in shared memory; a thread owns 2x2
// entries of the lower triangle for ALL C output channels, forms each lag product once and scales it by the C weights
// (4 + 2C FMA per entry-frame instead of 4C).  The right-hand side rides along as row L of the matrix (the Cholesky factor
// of [[R, r], [r^H, .]] carries L^-1 r in its last row), so the forward substitution costs nothing extra; k_wpe_chol is a
// right-looking panel Cholesky (panel in shared memory, trailing update on the L2-resident workspace) followed by a
// panel-wise backward substitution; in fp64 its trailing update runs on the fp64 tensor cores (mma.sync m8n8k4, SASS DMMA) and the
// next diagonal block is factored ahead by one warp while the others finish the update (see k_wpe_chol).  Arithmetic: fp64 like the
// reference (RT = double; the loaded normal equations need it for the 1e-4 parity budget) or fp32 (cfg.wpe.fp32_normal_equations) on
// the CUDA cores; DESIGN.md K7 says why the Gram is not a tensor-core GEMM.
//
// Two forms of the same normal equations.  With A = [lags(s)] (L x S, S = estimation frames - lower), Theta_c = diag(theta_c) and
// ybar_c(s) = conj(x_c(s)), calc_Rr_ + load_R_ build  (A Theta_c^-1 A^H + delta_c I) g_c = A Theta_c^-1 ybar_c  with the uniform
// loading  delta_c = bias + load_factor (max_i (A Theta_c^-1 A^H)_ii + bias).  By the push-through identity
//     g_c = A (K + delta_c Theta_c)^-1 ybar_c,      K = A^H A   (S x S, the same for every channel and every iteration),
// so when an utterance is shorter than the filter (S < L: 5 s at D = 512 are 157 frames against L = 8 x 33 = 264) the batch is
// served in this FRAME-DOMAIN form: k_wpe_gram_dual forms K once per problem, k_wpe_chol<DUAL> adds delta_c Theta_c on the fly,
// factors the S x S system and maps the solution back through A.  Symmetric diagonal scaling (Theta^1/2) turns K + delta Theta
// into delta I + Theta^-1/2 K Theta^-1/2, which has the spectrum of the lag-domain matrix, and Cholesky's backward error does not
// depend on such scaling: same conditioning, (S / L)^3 of the factorisation work, and the Gram is not repeated per channel or
// iteration.  a.form picks the form per batch (btkb_api.cu: do_wpe).
#include "btkb_internal.h"
#include <math.h>
#include <algorithm>
#include <type_traits>

namespace btkb {
namespace {

constexpr int WPE_NB = 16;        // Cholesky panel width
constexpr int WPE_CORR_THREADS = 256;
constexpr int WPE_CHOL_THREADS = 256;
constexpr int WPE_CHOL_THREADS_FRAME = 128;   // frame-domain form: smaller systems, four CTAs per SM

template <typename RT> struct cx { RT x, y; };
template <typename RT> __device__ __forceinline__ cx<RT> mk(RT x, RT y) { cx<RT> r; r.x = x; r.y = y; return r; }
template <typename RT> __device__ __forceinline__ cx<RT> cmulc(cx<RT> a, cx<RT> b) { return mk<RT>(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y)); }  // a conj(b)
template <typename RT> __device__ __forceinline__ void cmsubc(cx<RT>& s, cx<RT> a, cx<RT> b) {  // s -= a conj(b)
  s.x = fma(-a.x, b.x, s.x); s.x = fma(-a.y, b.y, s.x); s.y = fma(-a.y, b.x, s.y); s.y = fma(a.x, b.y, s.y);
}
__device__ __forceinline__ void cmacf(float2& s, float2 a, float2 b) { s.x = fmaf(a.x, b.x, s.x); s.x = fmaf(-a.y, b.y, s.x); s.y = fmaf(a.x, b.y, s.y); s.y = fmaf(a.y, b.x, s.y); }

// problem index q -> chain g = u*K + k, k over the estimated band 0..nbins-1 (dereverberation.cc:669: bins inside
// (lower_bandWidthN_, upper_bandWidthN_) are skipped; of the bins the output stage reads, those are k > lower_bandWidthN_)
__device__ __forceinline__ int problem_chain(const WpeArgs& a, int q) { const int u = q / a.nbins; return u * a.K + (q - u * a.nbins); }
__device__ __forceinline__ int est_frames_of(const WpeArgs& a, int u) {
  const int Tu = frames_of(a.lengths[u], a.D, a.laN, a.pdA);
  return (a.est_frames >= 0) ? min(Tu, a.est_frames) : Tu;
}

__global__ void k_wpe_gather(WpeArgs a) {
  __shared__ float2 tile[32][33];
  const int c = blockIdx.y;
  const int g0 = blockIdx.x * 32, t0 = blockIdx.z * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, g = g0 + threadIdx.x;
    tile[j][threadIdx.x] = (t < a.T && g < a.G) ? a.X[((size_t)t * a.C + c) * a.Gp + g] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int g = g0 + j, t = t0 + threadIdx.x;
    if (g < a.G && t < a.Ts) a.S[((size_t)g * a.C + c) * a.Ts + t] = (t < a.T) ? tile[threadIdx.x][j] : make_float2(0.f, 0.f);
  }
}

// shared-memory series with P leading zeros: xs[c][P + t]
__device__ __forceinline__ void load_series(const WpeArgs& a, int g, int nfr, float2* xs, int xstride) {
  for (int i = threadIdx.x; i < a.C * xstride; i += blockDim.x) {
    const int c = i / xstride, j = i - c * xstride, t = j - a.P;
    xs[i] = (t >= 0 && t < nfr) ? a.S[((size_t)g * a.C + c) * a.Ts + t] : make_float2(0.f, 0.f);
  }
}
template <typename RT>
__device__ __forceinline__ void load_series_t(const WpeArgs& a, int g, int nfr, cx<RT>* xs, int xstride) {
  for (int i = threadIdx.x; i < a.C * xstride; i += blockDim.x) {
    const int c = i / xstride, j = i - c * xstride, t = j - a.P;
    const float2 v = (t >= 0 && t < nfr) ? a.S[((size_t)g * a.C + c) * a.Ts + t] : make_float2(0.f, 0.f);
    xs[i] = mk<RT>((RT)v.x, (RT)v.y);
  }
}

// MODE 0: theta over the estimation frames; MODE 1: the output stage over every frame of the utterance (writes X in place)
template <int MODE>
__global__ void __launch_bounds__(128) k_wpe_resid(WpeArgs a, int q0) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int g = (MODE == 0) ? problem_chain(a, q0 + blockIdx.x) : blockIdx.x;
  if (g >= a.G) return;
  const int u = g / a.K, k = g - u * a.K;
  const int Tu = frames_of(a.lengths[u], a.D, a.laN, a.pdA);
  const int nfr = (MODE == 0) ? est_frames_of(a, u) : Tu;
  const bool in_band = k < a.nbins;
  if (MODE == 1 && !in_band) return;  // X already holds the unprocessed value
  const int xstride = a.P + a.T;
  float2* xs = reinterpret_cast<float2*>(smem);
  float2* Gs = xs + (size_t)a.C * xstride;   // conj(G) [C][L]
  load_series(a, g, nfr, xs, xstride);
  for (int i = threadIdx.x; i < a.C * a.L; i += blockDim.x) { const float2 v = a.Gf[(size_t)g * a.C * a.L + i]; Gs[i] = make_float2(v.x, -v.y); }
  __syncthreads();
  // the sliding buffer of the output stage holds P frames: lags beyond P-1-lower read as zero (dereverberation.cc:473-490)
  const int lmax = (MODE == 1) ? min(a.P - 1, a.P - 1 - a.lowerN) : a.P - 1;
  for (int idx = threadIdx.x; idx < a.C * nfr; idx += blockDim.x) {
    const int c = idx / nfr, t = idx - c * nfr;
    const float2 x = xs[c * xstride + a.P + t];
    float2 acc = make_float2(0.f, 0.f);
    if (t >= a.lowerN) {
      const float2* gr = Gs + (size_t)c * a.L;
      for (int cp = 0; cp < a.C; cp++) {
        const float2* xr = xs + cp * xstride + a.P + t - a.lowerN;
        const float2* gc = gr + cp * a.P;
        for (int l = 0; l <= lmax; l++) cmacf(acc, gc[l], xr[-l]);
      }
    }
    const float2 d = make_float2(x.x - acc.x, x.y - acc.y);
    if (MODE == 0) {
      float th = sqrtf(fmaf(d.x, d.x, d.y * d.y));
      th = fmaxf(th, 1.0e-3f);  // subband_floor_ (dereverberation.cc:617)
      a.TH[((size_t)g * a.C + c) * a.Ts + t] = th * th;
    } else {
      a.X[((size_t)t * a.C + c) * a.Gp + g] = d;
    }
  }
}

// R_c (lower triangle) and the augmented row L = conj(r_c) for the problems [q0, q0 + gridDim.y); blockIdx.x = split.
// RT = double reproduces the reference's fp64 normal equations (light diagonal loading leaves lambda_min / lambda_max near
// 1e-5, below what an fp32 running sum of T terms resolves); RT = float is the fast variant for well-loaded problems.
template <int C, typename RT>
__global__ void __launch_bounds__(WPE_CORR_THREADS) k_wpe_corr(WpeArgs a, int q0) {
  extern __shared__ __align__(16) unsigned char smem[];
  typedef cx<RT> CX;
  const int q = q0 + blockIdx.y;
  const int g = problem_chain(a, q);
  const int u = g / a.K;
  const int nfr = est_frames_of(a, u);
  const int xstride = a.P + a.T;
  CX* xs = reinterpret_cast<CX*>(smem);
  RT* ws = reinterpret_cast<RT*>(xs + (size_t)C * xstride);  // [T][C] weights 1/theta, 0 outside the estimation frames
  load_series_t<RT>(a, g, nfr, xs, xstride);
  for (int i = threadIdx.x; i < C * a.T; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    ws[i] = (t < nfr && t >= a.lowerN) ? (RT)1 / (RT)a.TH[((size_t)g * C + c) * a.Ts + t] : (RT)0;
  }
  __syncthreads();
  const int L = a.L, P = a.P, Lr = a.Lr;
  CX* Rq = reinterpret_cast<CX*>(a.Rw) + (size_t)blockIdx.y * C * a.slot;
  const int nS = max(nfr - a.lowerN, 0);   // s' = s - lower, s = lower .. nfr-1
  const int L2 = (L + 1) / 2;
  const int ntiles = L2 * (L2 + 1) / 2;
  for (int tq = blockIdx.x * blockDim.x + threadIdx.x; tq < ntiles; tq += gridDim.x * blockDim.x) {
    int bi = (int)((sqrtf(8.0f * (float)tq + 1.0f) - 1.0f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= tq) bi++;
    while (bi * (bi + 1) / 2 > tq) bi--;
    const int bj = tq - bi * (bi + 1) / 2;
    const int i0 = 2 * bi, i1 = min(i0 + 1, L - 1), j0 = 2 * bj, j1 = min(j0 + 1, L - 1);
    const CX* pi0 = xs + (i0 / P) * xstride + P - (i0 % P);
    const CX* pi1 = xs + (i1 / P) * xstride + P - (i1 % P);
    const CX* pj0 = xs + (j0 / P) * xstride + P - (j0 % P);
    const CX* pj1 = xs + (j1 / P) * xstride + P - (j1 % P);
    CX acc[4][C];
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
      for (int c = 0; c < C; c++) acc[e][c] = mk<RT>(0, 0);
    for (int s = 0; s < nS; s++) {
      const CX a0 = pi0[s], a1 = pi1[s], b0 = pj0[s], b1 = pj1[s];
      const CX p00 = cmulc(a0, b0), p01 = cmulc(a0, b1), p10 = cmulc(a1, b0), p11 = cmulc(a1, b1);
      const RT* w = ws + (size_t)(s + a.lowerN) * C;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const RT wc = w[c];
        acc[0][c].x = fma(wc, p00.x, acc[0][c].x); acc[0][c].y = fma(wc, p00.y, acc[0][c].y);
        acc[1][c].x = fma(wc, p01.x, acc[1][c].x); acc[1][c].y = fma(wc, p01.y, acc[1][c].y);
        acc[2][c].x = fma(wc, p10.x, acc[2][c].x); acc[2][c].y = fma(wc, p10.y, acc[2][c].y);
        acc[3][c].x = fma(wc, p11.x, acc[3][c].x); acc[3][c].y = fma(wc, p11.y, acc[3][c].y);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      CX* Rc = Rq + (size_t)c * a.slot;
      Rc[(size_t)i0 * Lr + j0] = acc[0][c];
      if (j0 + 1 < L && j0 + 1 <= i0) Rc[(size_t)i0 * Lr + j0 + 1] = acc[1][c];
      if (i0 + 1 < L) {
        Rc[(size_t)(i0 + 1) * Lr + j0] = acc[2][c];
        if (j0 + 1 < L) Rc[(size_t)(i0 + 1) * Lr + j0 + 1] = acc[3][c];
      }
    }
  }
  // augmented row: A_c[L][j] = conj(r_c[j]) = sum_s x_c(s) conj(lags_j(s)) / theta_c(s)
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
    const CX* pj = xs + (j / P) * xstride + P - (j % P);
    CX acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = mk<RT>(0, 0);
    for (int s = 0; s < nS; s++) {
      const CX b = pj[s];
      const RT* w = ws + (size_t)(s + a.lowerN) * C;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const CX x = xs[c * xstride + P + s + a.lowerN];
        const CX pr = cmulc(x, b);
        acc[c].x = fma(w[c], pr.x, acc[c].x); acc[c].y = fma(w[c], pr.y, acc[c].y);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) Rq[(size_t)c * a.slot + (size_t)L * Lr + j] = acc[c];
  }
}

// Frame-domain form: K[s][s'] = sum_i conj(lags_i(s)) lags_i(s') (lower triangle, s' <= s) of the problems [q0, q0 + gridDim.x) into
// slot C of each problem.  With lags_(c, l)(s) = x_c(s - l) the entry on diagonal d = s - s' is a window sum of ONE sequence,
//     K[s' + d][s'] = sum_{l < P} p_d(s' - l),      p_d(t) = sum_c conj(x_c(t + d)) x_c(t)   (0 for t < 0),
// so a diagonal costs C complex MACs per entry for p_d and P additions for the window instead of C P MACs — and the window is summed
// directly (no running add / subtract: after a loud passage that would leave rounding residue where K is exactly small).  One CTA per
// problem, WPE_GRAM_DG diagonals at a time through a shared-memory line buffer with P - 1 leading zeros.
constexpr int WPE_GRAM_DG = 8;
template <typename RT>
__global__ void __launch_bounds__(WPE_CORR_THREADS) k_wpe_gram_dual(WpeArgs a, int q0) {
  extern __shared__ __align__(16) unsigned char smem[];
  typedef cx<RT> CX;
  const int g = problem_chain(a, q0 + blockIdx.x);
  const int u = g / a.K;
  const int nfr = est_frames_of(a, u);
  const int xstride = a.P + a.T, P = a.P, C = a.C, Lr = a.Lr;
  const int pstride = a.Sd + P;                       // line buffer: [P - 1 zeros | p_d(0 .. nS - d - 1)]
  CX* xs = reinterpret_cast<CX*>(smem);
  CX* pb = xs + (size_t)C * xstride;                  // [WPE_GRAM_DG][pstride]
  load_series_t<RT>(a, g, nfr, xs, xstride);
  for (int i = threadIdx.x; i < WPE_GRAM_DG * pstride; i += blockDim.x) pb[i] = mk<RT>(0, 0);
  __syncthreads();
  const int nS = max(nfr - a.lowerN, 0);
  CX* Kq = reinterpret_cast<CX*>(a.Rw) + ((size_t)blockIdx.x * (C + 1) + C) * a.slot;
  for (int d0 = 0; d0 < nS; d0 += WPE_GRAM_DG) {
    for (int i = threadIdx.x; i < WPE_GRAM_DG * nS; i += blockDim.x) {
      const int dd = i / nS, t = i - dd * nS, d = d0 + dd;
      CX acc = mk<RT>(0, 0);
      if (t + d < nS) {
        for (int c = 0; c < C; c++) {
          const CX av = xs[(size_t)c * xstride + P + t + d], b = xs[(size_t)c * xstride + P + t];   // acc += b conj(av)
          acc.x = fma(b.x, av.x, acc.x); acc.x = fma(b.y, av.y, acc.x); acc.y = fma(b.y, av.x, acc.y); acc.y = fma(-b.x, av.y, acc.y);
        }
      }
      pb[(size_t)dd * pstride + P - 1 + t] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < WPE_GRAM_DG * nS; i += blockDim.x) {
      const int dd = i / nS, t = i - dd * nS, d = d0 + dd;
      if (t + d < nS) {
        const CX* pp = pb + (size_t)dd * pstride + P - 1 + t;
        CX e0 = mk<RT>(0, 0), e1 = e0;
        int l = 0;
        for (; l + 1 < P; l += 2) { const CX v0 = pp[-l], v1 = pp[-l - 1]; e0.x += v0.x; e0.y += v0.y; e1.x += v1.x; e1.y += v1.y; }
        if (l < P) { const CX v0 = pp[-l]; e0.x += v0.x; e0.y += v0.y; }
        CX k = mk<RT>(e0.x + e1.x, e0.y + e1.y);
        if (d == 0) k.y = 0;
        Kq[(size_t)(t + d) * Lr + t] = k;
      }
    }
    __syncthreads();
  }
}

// ---- k_wpe_chol: shared-memory layout
// The panel [rows][WPE_NB] lives in shared memory as two planes (real, imaginary) of RT with a row stride of WPE_PS = 20 elements
// and the column index XOR-swizzled with the low four bits of the row: idx(r, j) = r * 20 + (j ^ (r & 15)).  That one layout is
// conflict-free for both access shapes of the kernel: "every lane its own row, all lanes the same column" (scaling, TRSM, panel
// load / store: 16 rows x one column hit 16 different 8-byte bank pairs, because the swizzle spreads a column over the 16
// positions of a row group and the stride rotates the groups), and the fp64 tensor-core fragments of the trailing update (lane
// l reads row l / 4, column k0 + l % 4: four rows x four columns per half-warp = 16 different bank pairs, because 20 elements
// = 8 banks mod 32 separate the rows and the swizzle only permutes the four columns inside their aligned group).
// RT = float keeps the interleaved layout (one 8-byte access per complex value, row stride 17 values): it has no tensor-core path.
constexpr int WPE_PS = 20;
constexpr int WPE_LD = WPE_NB + 1;
__device__ __forceinline__ int pidx(int r, int j) { return r * WPE_PS + (j ^ (r & 15)); }
// RT elements of the panel buffer: two planes of Lcap + 1 rows; (frame-domain form) at least the nseries float2 series values
// that alias it outside the factorisation
template <typename RT>
__host__ __device__ inline size_t chol_region0(int Lcap, int nseries) {
  size_t e = (size_t)2 * (Lcap + 1) * (sizeof(RT) == 8 ? WPE_PS : WPE_LD);
  const size_t e3 = ((size_t)nseries * sizeof(float2) + sizeof(RT) - 1) / sizeof(RT);
  if (e3 > e) e = e3;
  return (e + 1) & ~(size_t)1;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// one CTA per (problem, channel): diagonal bias + loading, panel Cholesky of the augmented matrix, backward substitution.
// DUAL: the frame-domain system (K + delta_c Theta_c) z = ybar_c of size S = the utterance's estimation frames - lower, read from the
// problem's shared K while the first panel is processed, then g_c = A z.
// MMA (RT = double): the trailing update A22 -= L21 L21^H runs on the fp64 tensor cores (mma.sync m8n8k4: measured 37.1 TFLOP/s on
// this GPU, the rate of the DFMA pipe, with an eighth of the issue slots and a third of the shared-memory traffic of the scalar
// tile): a warp owns a 16 x 16 block of the lower triangle, 2 x 2 fragments, real and imaginary parts as four real products.
template <typename RT, bool DUAL, bool MMA>
__global__ void __launch_bounds__(DUAL ? WPE_CHOL_THREADS_FRAME : WPE_CHOL_THREADS, DUAL ? 4 : 2) k_wpe_chol(WpeArgs a, int q0, int C) {
  extern __shared__ __align__(16) unsigned char smem[];
  typedef cx<RT> CX;
  const int qq = blockIdx.x / C, c = blockIdx.x - qq * C;
  const int g = problem_chain(a, q0 + qq);
  const int Lcap = DUAL ? a.Sd : a.L;     // sizes the shared-memory layout (the same for every CTA of the launch)
  const int nfr = DUAL ? est_frames_of(a, g / a.K) : 0;
  const int L = DUAL ? max(nfr - a.lowerN, 0) : a.L;
  const int Lr = a.Lr, n = L + 1;
  CX* A = reinterpret_cast<CX*>(a.Rw) + ((size_t)qq * (DUAL ? C + 1 : C) + c) * a.slot;
  const CX* Kq = reinterpret_cast<const CX*>(a.Rw) + ((size_t)qq * (C + 1) + C) * a.slot;   // DUAL only
  const int xstride = a.P + a.T;
  RT* Pre = reinterpret_cast<RT*>(smem);                   // current panel, rows relative to j0: real plane ...
  RT* Pim = Pre + (size_t)(Lcap + 1) * WPE_PS;             // ... and imaginary plane
  RT* red = Pre + chol_region0<RT>(Lcap, DUAL ? C * xstride : 0);   // [16] block reduction, [16] 1 / L_jj of the current diagonal block
  RT* dinv = red + 16;
  unsigned char* tri = reinterpret_cast<unsigned char*>(red + 32);   // [136][2] (row, column) of the e-th entry of a lower triangle
  CX* yv = reinterpret_cast<CX*>(tri + 272);               // [Lcap] back-substitution vector (DUAL prologue: 1 / theta as scratch)
  RT* th = reinterpret_cast<RT*>(yv + Lcap);               // DUAL: [Lcap] theta_c(s + lower)
  float2* xs = reinterpret_cast<float2*>(smem);            // DUAL: the C series [C][P + T]; ALIASES the panel buffer — alive only before
                                                           // the first panel is loaded (delta) and after the back substitution (g = A z)
  constexpr bool PLANAR = sizeof(RT) == 8;
  CX* Pn = reinterpret_cast<CX*>(smem);                    // RT = float: interleaved panel
  auto ldp = [&](int r, int j) -> CX { if (PLANAR) { const int i = pidx(r, j); return mk<RT>(Pre[i], Pim[i]); } return Pn[r * WPE_LD + j]; };
  auto stp = [&](int r, int j, const CX& v) { if (PLANAR) { const int i = pidx(r, j); Pre[i] = v.x; Pim[i] = v.y; } else Pn[r * WPE_LD + j] = v; };
  const int tid = threadIdx.x;
  const RT bias = (RT)a.diagonal_bias, loadf = (RT)a.load_factor;
  RT delta = 0;
  for (int e = tid; e < 136; e += blockDim.x) {
    int rr = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((rr + 1) * (rr + 2) / 2 <= e) rr++;
    while (rr * (rr + 1) / 2 > e) rr--;
    tri[2 * e] = (unsigned char)rr; tri[2 * e + 1] = (unsigned char)(e - rr * (rr + 1) / 2);
  }

  if (DUAL) {
    // ---- delta_c = bias + load_factor (max_i R_ii + bias), R_ii = sum_s |lags_i(s)|^2 / theta_c(s)   (calc_Rr_ :577-580, load_R_ :648-663)
    load_series(a, g, nfr, xs, xstride);
    RT* wi = reinterpret_cast<RT*>(yv);                    // 1 / theta
    for (int s2 = tid; s2 < L; s2 += blockDim.x) {
      const RT t = (RT)a.TH[((size_t)g * C + c) * a.Ts + s2 + a.lowerN];
      th[s2] = t; wi[s2] = (RT)1 / t;
    }
    __syncthreads();
    RT mx = 0;
    for (int i = tid; i < a.L; i += blockDim.x) {
      const float2* pi = xs + (size_t)(i / a.P) * xstride + a.P - (i % a.P);
      RT acc = 0;
      for (int s2 = 0; s2 < L; s2++) { const float2 v = pi[s2]; acc = fma(fma((RT)v.x, (RT)v.x, (RT)v.y * (RT)v.y), wi[s2], acc); }
      mx = fmax(mx, acc + bias);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, red[w]);
    delta = bias + mx * loadf;
    __syncthreads();                                       // the series are dead: the panel buffer may be overwritten
  } else {
    // ---- diagonal: + diagonal_bias (calc_Rr_, :577-580), then |d| + max|d| load_factor (load_R_, :648-663)
    RT mx = 0;
    for (int i = tid; i < L; i += blockDim.x) {
      CX d = A[(size_t)i * Lr + i];
      d.x += bias;
      mx = fmax(mx, sqrt(fma(d.x, d.x, d.y * d.y)));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, red[w]);
    for (int i = tid; i < L; i += blockDim.x) {
      CX d = A[(size_t)i * Lr + i];
      d.x += bias;
      A[(size_t)i * Lr + i] = mk<RT>(sqrt(fma(d.x, d.x, d.y * d.y)) + mx * loadf, 0);
    }
    __syncthreads();
  }
  // entry (row, col), col <= row, of the system before any update: DUAL reads K, the loading and the right-hand side; row L is the
  // augmented row conj(rhs) = x_c(s + lower)
  auto src0 = [&](int row, int col) -> CX {
    if (!DUAL) return A[(size_t)row * Lr + col];
    if (row == L) {
      if (col == L) return mk<RT>(0, 0);
      const float2 v = a.S[((size_t)g * C + c) * a.Ts + col + a.lowerN];
      return mk<RT>((RT)v.x, (RT)v.y);
    }
    CX v = Kq[(size_t)row * Lr + col];
    if (row == col) { v.x = fma(delta, th[row], v.x); v.y = 0; }
    return v;
  };

  bool bad = false;
  bool ahead = false;   // look-ahead (tensor-core path): rows 0 .. 15 of the panel buffer already hold this panel's FACTORED diagonal block
  int* ctr = reinterpret_cast<int*>(red + 15);   // next unclaimed block of the trailing update
  for (int j0 = 0; j0 < L; j0 += WPE_NB) {
    const int nb = min(WPE_NB, L - j0);
    const int nrows = n - j0;   // rows j0 .. L (the augmented row included)
    // ---- load the panel (columns >= nb and the part above the diagonal as zeros)
    // (four loads in flight per thread before the first store: the values come from L2)
    for (int i0 = tid; i0 < nrows * WPE_NB; i0 += 4 * blockDim.x) {
      CX v4[4];
#pragma unroll
      for (int u4 = 0; u4 < 4; u4++) {
        const int i = i0 + u4 * blockDim.x, r = i / WPE_NB, jj = i - r * WPE_NB;
        v4[u4] = (i < nrows * WPE_NB && jj < nb && jj <= r && !(ahead && r < WPE_NB)) ? ((DUAL && j0 == 0) ? src0(r, jj) : A[(size_t)(j0 + r) * Lr + j0 + jj]) : mk<RT>(0, 0);
      }
#pragma unroll
      for (int u4 = 0; u4 < 4; u4++) {
        const int i = i0 + u4 * blockDim.x, r = i / WPE_NB, jj = i - r * WPE_NB;
        if (i < nrows * WPE_NB && !(ahead && r < WPE_NB)) stp(r, jj, v4[u4]);
      }
    }
    __syncthreads();
    // ---- factor the nb x nb diagonal block with ONE warp (no CTA barrier inside): per column a scaling by 1 / sqrt(pivot), then the
    // rank-one update of the remaining triangle with its (row, column) pairs dealt out to the 32 lanes
    auto factor_diag = [&](int nbd) {   // warp 0, all 32 lanes
      for (int jj = 0; jj < nbd; jj++) {
        const RT dj = ldp(jj, jj).x;
        if (!(dj > (RT)0)) bad = true;
        const RT inv = rsqrt(fmax(dj, (RT)1e-30));
        __syncwarp();   // every lane has read the pivot before its owner overwrites it
        if (tid >= jj && tid < nbd) {
          if (tid == jj) { stp(tid, jj, mk<RT>(dj * inv, 0)); dinv[jj] = inv; }
          else { const CX v = ldp(tid, jj); stp(tid, jj, mk<RT>(v.x * inv, v.y * inv)); }
        }
        __syncwarp();
        const int m = nbd - 1 - jj;
        for (int e = tid; e < m * (m + 1) / 2; e += 32) {
          const int r = jj + 1 + tri[2 * e], kk = jj + 1 + tri[2 * e + 1];
          CX v = ldp(r, kk);
          cmsubc(v, ldp(r, jj), ldp(kk, jj));
          if (kk == r) v.y = 0;
          stp(r, kk, v);
        }
        __syncwarp();
      }
    };
    if (!ahead) {
      if (tid < 32) factor_diag(nb);
      __syncthreads();
    }
    // the factored diagonal block goes back to A here (the backward substitution reads it): its rows of the panel buffer are free
    // from the next barrier on
    for (int i = tid; i < nb * WPE_NB; i += blockDim.x) {
      const int r = i / WPE_NB, jj = i - r * WPE_NB;
      if (jj < nb && jj <= r) A[(size_t)(j0 + r) * Lr + j0 + jj] = ldp(r, jj);
    }
    // ---- ... then the rows below it are independent: L21 = A21 L11^-H, one thread per row held in registers, right-looking (once
    // x_jj is final the remaining entries of the row take their updates independently of one another), no barrier
    for (int r = nb + tid; r < nrows; r += blockDim.x) {
      CX v[WPE_NB];
#pragma unroll
      for (int jj = 0; jj < WPE_NB; jj++) v[jj] = ldp(r, jj);
#pragma unroll
      for (int jj = 0; jj < WPE_NB; jj++) {
        if (jj < nb) {
          const RT inv = dinv[jj];
          v[jj] = mk<RT>(v[jj].x * inv, v[jj].y * inv);
#pragma unroll
          for (int q = jj + 1; q < WPE_NB; q++)
            if (q < nb) cmsubc(v[q], v[jj], ldp(q, jj));
          stp(r, jj, v[jj]);
        }
      }
    }
    // ---- trailing update: A[i][k] -= sum_jj Pn[i][jj] conj(Pn[k][jj]) for j1 <= k <= i <= L
    const int j1 = j0 + nb;
    const int nt = n - j1;             // trailing rows (incl. the augmented one)
    // look-ahead: when the next panel is a full one, warp 0 takes block (0, 0) of the trailing triangle — the next diagonal block —
    // first, keeps its updated values in rows 0 .. 15 of the panel buffer instead of storing them, and factors them there while the
    // other warps work through the remaining blocks (claimed from a shared counter, so that warp 0 simply joins in late)
    const bool look = MMA && (L - j1 >= WPE_NB);
    if (tid == 0) *ctr = look ? 1 : 0;
    __syncthreads();
    // ---- write the factored rows below the diagonal block back (needed by the backward substitution)
    for (int i = nb * WPE_NB + tid; i < nrows * WPE_NB; i += blockDim.x) {
      const int r = i / WPE_NB, jj = i - r * WPE_NB;
      if (jj < nb) A[(size_t)(j0 + r) * Lr + j0 + jj] = ldp(r, jj);
    }
    if (MMA && nt > 0) {
      // A warp owns 16 x 16 blocks (bi >= bk) of the trailing triangle.  S = L21(bi) L21(bk)^H as four real products per fragment:
      // Sr = Ar Br^T + Ai Bi^T, Si = Ai Br^T - Ar Bi^T; the fragment of B^T is read from the panel exactly like the one of A (lane l:
      // row l / 4, column k0 + l % 4).  A lane ends up with rows bi 16 + ti 8 + l / 4 and the column pairs bk 16 + tk 8 + 2 (l % 4) + {0, 1}:
      // 32 contiguous bytes per read-modify-write of A.  Rows past the end are clamped for the loads and dropped at the store; columns
      // >= nb of the panel are zero.
      const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
      const int nb16 = (nt + 15) >> 4;
      bool mine = look && warp == 0;   // block (0, 0) is still to be done by this warp
      for (;;) {
        int bp = 0;
        if (!mine) {
          if (lane == 0) bp = atomicAdd(ctr, 1);
          bp = __shfl_sync(0xffffffffu, bp, 0);
          if (bp >= nb16 * (nb16 + 1) / 2) break;
        }
        const bool keep = mine;          // this block stays on chip
        mine = false;
        int bi = (int)((sqrtf(8.0f * (float)bp + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= bp) bi++;
        while (bi * (bi + 1) / 2 > bp) bi--;
        const int bk = bp - bi * (bi + 1) / 2;
        const CX* base = (DUAL && j0 == 0) ? Kq : A;
        if (a.prefetch) {
#pragma unroll
          for (int ti = 0; ti < 2; ti++)
#pragma unroll
            for (int tk = 0; tk < 2; tk++) {
              const int row = bi * 16 + ti * 8 + lr, col = bk * 16 + tk * 8 + 2 * lc;
              if (row < nt && col < nt && col <= row) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)(j1 + row) * Lr + j1 + col));
            }
        }
        int ba[2], ca[2], bb[2], cb[2];   // pidx(r, k0 + lc) = r * WPE_PS + ((lc ^ (r & 15)) ^ k0) for k0 = 0, 4, 8, 12
#pragma unroll
        for (int t2 = 0; t2 < 2; t2++) {
          const int ra = nb + min(bi * 16 + t2 * 8 + lr, nt - 1), rb = nb + min(bk * 16 + t2 * 8 + lr, nt - 1);
          ba[t2] = ra * WPE_PS; ca[t2] = lc ^ (ra & 15); bb[t2] = rb * WPE_PS; cb[t2] = lc ^ (rb & 15);
        }
        double sr[2][2][2], si[2][2][2];
#pragma unroll
        for (int ti = 0; ti < 2; ti++)
#pragma unroll
          for (int tk = 0; tk < 2; tk++) { sr[ti][tk][0] = sr[ti][tk][1] = 0.0; si[ti][tk][0] = si[ti][tk][1] = 0.0; }
#pragma unroll
        for (int k0 = 0; k0 < WPE_NB; k0 += 4) {
          double ar[2], ai[2], br[2], bim[2], nbi[2];
#pragma unroll
          for (int t2 = 0; t2 < 2; t2++) {
            const int oa = ba[t2] + (ca[t2] ^ k0), ob = bb[t2] + (cb[t2] ^ k0);
            ar[t2] = (double)Pre[oa]; ai[t2] = (double)Pim[oa]; br[t2] = (double)Pre[ob]; bim[t2] = (double)Pim[ob]; nbi[t2] = -bim[t2];
          }
#pragma unroll
          for (int ti = 0; ti < 2; ti++)
#pragma unroll
            for (int tk = 0; tk < 2; tk++) {
              dmma884(sr[ti][tk][0], sr[ti][tk][1], ar[ti], br[tk]);
              dmma884(sr[ti][tk][0], sr[ti][tk][1], ai[ti], bim[tk]);
              dmma884(si[ti][tk][0], si[ti][tk][1], ai[ti], br[tk]);
              dmma884(si[ti][tk][0], si[ti][tk][1], ar[ti], nbi[tk]);
            }
        }
        // read-modify-write of the block: all eight old values are requested before the first store (the compiler cannot move a load
        // of A across a store to A, and each would otherwise be its own round trip to L2)
        CX old[2][2][2];
#pragma unroll
        for (int ti = 0; ti < 2; ti++)
#pragma unroll
          for (int tk = 0; tk < 2; tk++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int row = bi * 16 + ti * 8 + lr, col = bk * 16 + tk * 8 + 2 * lc + e;
              if (row < nt && col < nt && col <= row) old[ti][tk][e] = (DUAL && j0 == 0) ? src0(j1 + row, j1 + col) : A[(size_t)(j1 + row) * Lr + j1 + col];
            }
#pragma unroll
        for (int ti = 0; ti < 2; ti++)
#pragma unroll
          for (int tk = 0; tk < 2; tk++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int row = bi * 16 + ti * 8 + lr, col = bk * 16 + tk * 8 + 2 * lc + e;
              if (row < nt && col < nt && col <= row) {
                const CX v = mk<RT>(old[ti][tk][e].x - (RT)sr[ti][tk][e], old[ti][tk][e].y - (RT)si[ti][tk][e]);
                if (keep) stp(row, col, v); else A[(size_t)(j1 + row) * Lr + j1 + col] = v;
              }
            }
        if (keep) { __syncwarp(); factor_diag(WPE_NB); }
      }

    } else if (nt > 0) {
      // A thread owns the entries (ri + q nt4, rk + p nt4), q, p = 0..3, of the trailing block (nt4 = ceil(nt / 4)): the
      // four-way interleave puts the lanes of a warp on CONSECUTIVE panel rows and makes the global read-modify-write of A
      // coalesced along a row.  Of the 16 entries the 6 with q > p always lie in the lower triangle, the 4 with q == p do iff
      // rk <= ri, the rest never.
      const int nt4 = (nt + 3) / 4;
      for (int tq = tid; tq < nt4 * nt4; tq += blockDim.x) {
        const int ri = tq / nt4, rk = tq - ri * nt4;
        const bool diag = rk <= ri;
        if (a.prefetch) {
          // the entries this tile will read-modify-write come from L2 (or, first panel of the frame-domain form, from K): ask for them
          // now so that they sit in L1 when the 16-column product below is done — no registers are held while they travel
          const CX* base = (DUAL && j0 == 0) ? Kq : A;
#pragma unroll
          for (int q = 0; q < 4; q++)
#pragma unroll
            for (int pc = 0; pc <= q; pc++) {
              const int row = ri + q * nt4, col = rk + pc * nt4;
              if (row < nt && col < nt && (pc < q || diag))
                asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)(j1 + row) * Lr + j1 + col));
            }
        }
        int pa[4], pb[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { pa[q] = nb + min(ri + q * nt4, nt - 1); pb[q] = nb + min(rk + q * nt4, nt - 1); }
        CX sd[4], so[6];
#pragma unroll
        for (int e = 0; e < 4; e++) sd[e] = mk<RT>(0, 0);
#pragma unroll
        for (int e = 0; e < 6; e++) so[e] = mk<RT>(0, 0);
        for (int jj = 0; jj < nb; jj++) {
          CX av[4], bv[4];
#pragma unroll
          for (int q = 0; q < 4; q++) { av[q] = ldp(pa[q], jj); bv[q] = ldp(pb[q], jj); }
          if (diag) {
#pragma unroll
            for (int q = 0; q < 4; q++) cmsubc(sd[q], av[q], bv[q]);
          }
          cmsubc(so[0], av[1], bv[0]); cmsubc(so[1], av[2], bv[0]); cmsubc(so[2], av[2], bv[1]);
          cmsubc(so[3], av[3], bv[0]); cmsubc(so[4], av[3], bv[1]); cmsubc(so[5], av[3], bv[2]);
        }
        // all old values are requested before the first store (see the tensor-core branch)
        auto fetch = [&](int q, int pcol, CX& sv) {   // sv <- old + sv
          const int row = ri + q * nt4, col = rk + pcol * nt4;
          if (row < nt && col < nt) {
            const CX v = (DUAL && j0 == 0) ? src0(j1 + row, j1 + col) : A[(size_t)(j1 + row) * Lr + j1 + col];
            sv.x += v.x; sv.y += v.y;
          }
        };
        auto put = [&](int q, int pcol, const CX& sv) {
          const int row = ri + q * nt4, col = rk + pcol * nt4;
          if (row < nt && col < nt) A[(size_t)(j1 + row) * Lr + j1 + col] = sv;
        };
        if (diag) {
#pragma unroll
          for (int q = 0; q < 4; q++) fetch(q, q, sd[q]);
        }
        fetch(1, 0, so[0]); fetch(2, 0, so[1]); fetch(2, 1, so[2]); fetch(3, 0, so[3]); fetch(3, 1, so[4]); fetch(3, 2, so[5]);
        if (diag) {
#pragma unroll
          for (int q = 0; q < 4; q++) put(q, q, sd[q]);
        }
        put(1, 0, so[0]); put(2, 0, so[1]); put(2, 1, so[2]); put(3, 0, so[3]); put(3, 1, so[4]); put(3, 2, so[5]);
      }
    }
    ahead = look;
    __syncthreads();
  }
  if (bad && tid == 0) atomicExch(a.err_flag, 1);

  // ---- backward substitution L^H gvec = y, y_j = conj(A[L][j]); the diagonal block of a panel sits in rows 0 .. 15 of the panel buffer
  for (int j = tid; j < L; j += blockDim.x) { const CX v = A[(size_t)L * Lr + j]; yv[j] = mk<RT>(v.x, -v.y); }
  __syncthreads();
  for (int j0 = ((L - 1) / WPE_NB) * WPE_NB; j0 >= 0; j0 -= WPE_NB) {
    const int nb = min(WPE_NB, L - j0);
    for (int i = tid; i < nb * WPE_NB; i += blockDim.x) {
      const int r = i / WPE_NB, jj = i - r * WPE_NB;
      stp(r, jj, (jj <= r && jj < nb) ? A[(size_t)(j0 + r) * Lr + j0 + jj] : mk<RT>(0, 0));
    }
    __syncthreads();
    if (tid < 32) {   // column sweep: once g_jj is final, the lanes q < jj take y_q -= g_jj conj(L[jj][q])
      const RT invd = (tid < nb) ? (RT)1 / ldp(tid, tid).x : (RT)0;
      for (int jj = nb - 1; jj >= 0; jj--) {
        const CX yj = yv[j0 + jj];
        const RT inv = __shfl_sync(0xffffffffu, invd, jj);
        const CX gj = mk<RT>(yj.x * inv, yj.y * inv);
        __syncwarp();
        if (tid == jj) yv[j0 + jj] = gj;
        else if (tid < jj) { CX sv = yv[j0 + tid]; cmsubc(sv, gj, ldp(jj, tid)); yv[j0 + tid] = sv; }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int k = tid; k < j0; k += blockDim.x) {   // y_k -= sum_i conj(L[i][k]) g_i over the panel rows
      CX s = yv[k];
      for (int r = 0; r < nb; r++) cmsubc(s, yv[j0 + r], A[(size_t)(j0 + r) * Lr + k]);
      yv[k] = s;
    }
    __syncthreads();
  }
  if (DUAL) {   // g_c = A z: g_i = sum_s lags_i(s) z_s
    load_series(a, g, nfr, xs, xstride);
    __syncthreads();
    for (int i = tid; i < a.L; i += blockDim.x) {
      const float2* pi = xs + (size_t)(i / a.P) * xstride + a.P - (i % a.P);
      CX acc = mk<RT>(0, 0);
      for (int s2 = 0; s2 < L; s2++) {
        const float2 v = pi[s2]; const CX z = yv[s2];
        acc.x = fma((RT)v.x, z.x, acc.x); acc.x = fma(-(RT)v.y, z.y, acc.x); acc.y = fma((RT)v.x, z.y, acc.y); acc.y = fma((RT)v.y, z.x, acc.y);
      }
      a.Gf[((size_t)g * C + c) * a.L + i] = make_float2((float)acc.x, (float)acc.y);
    }
  } else {
    for (int j = tid; j < L; j += blockDim.x) a.Gf[((size_t)g * C + c) * L + j] = make_float2((float)yv[j].x, (float)yv[j].y);
  }
}

}  // namespace

size_t wpe_workspace_bytes(int C, size_t slot, int chunk, int fp32) {
  return (size_t)chunk * (C + 1) * slot * (fp32 ? sizeof(float2) : sizeof(double2));   // C systems + the shared K of the frame-domain form
}

template <typename RT>
static cudaError_t launch_wpe_t(const WpeArgs& a, int chunk, cudaStream_t st, int* launches) {
  const int C = a.C;
  cudaError_t e;
  {
    dim3 grid((a.G + 31) / 32, C, (a.Ts + 31) / 32), block(32, 8);
    k_wpe_gather<<<grid, block, 0, st>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    (*launches)++;
  }
  if (!a.apply_only) {
    e = cudaMemsetAsync(a.Gf, 0, (size_t)a.G * C * a.L * sizeof(float2), st);
    if (e != cudaSuccess) return e;
  }
  const bool dual = a.form == 1;
  const int xstride = a.P + a.T;
  const int Lcap = dual ? a.Sd : a.L;
  if ((size_t)(Lcap + 1) * a.Lr > a.slot || Lcap + 1 > a.Lr) return cudaErrorInvalidValue;
  const size_t sm_resid = ((size_t)C * xstride + (size_t)C * a.L) * sizeof(float2);
  const size_t sm_corr = (size_t)C * xstride * sizeof(cx<RT>) + (dual ? (size_t)WPE_GRAM_DG * (a.Sd + a.P) * sizeof(cx<RT>) : (size_t)C * a.T * sizeof(RT));
  const size_t sm_chol = chol_region0<RT>(Lcap, dual ? C * xstride : 0) * sizeof(RT) + 32 * sizeof(RT) + 272 + (size_t)Lcap * sizeof(cx<RT>) +
                         (dual ? (size_t)Lcap * sizeof(RT) : 0);
  const bool mma = std::is_same<RT, double>::value && a.mma;
  const int chol_threads = dual ? std::min(a.chol_threads > 0 ? a.chol_threads : WPE_CHOL_THREADS_FRAME, WPE_CHOL_THREADS_FRAME)
                                : (a.chol_threads > 0 ? a.chol_threads : WPE_CHOL_THREADS);
  if (dual) chunk = a.chunk_frame > 0 ? a.chunk_frame : chunk;
  if (sm_resid > 200 * 1024 || sm_corr > 200 * 1024 || sm_chol > 200 * 1024) return cudaErrorInvalidValue;
  if ((e = cudaFuncSetAttribute(k_wpe_resid<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_resid)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_wpe_resid<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_resid)) != cudaSuccess) return e;
  void (*chol)(WpeArgs, int, int) = dual ? (mma ? k_wpe_chol<RT, true, true> : k_wpe_chol<RT, true, false>) : (mma ? k_wpe_chol<RT, false, true> : k_wpe_chol<RT, false, false>);
  if ((e = cudaFuncSetAttribute(chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chol)) != cudaSuccess) return e;
  void (*corr)(WpeArgs, int) = nullptr;
  if (dual) corr = k_wpe_gram_dual<RT>;
  else switch (C) {
    case 1: corr = k_wpe_corr<1, RT>; break;
    case 2: corr = k_wpe_corr<2, RT>; break;
    case 3: corr = k_wpe_corr<3, RT>; break;
    case 4: corr = k_wpe_corr<4, RT>; break;
    case 5: corr = k_wpe_corr<5, RT>; break;
    case 6: corr = k_wpe_corr<6, RT>; break;
    case 7: corr = k_wpe_corr<7, RT>; break;
    case 8: corr = k_wpe_corr<8, RT>; break;
    default: return cudaErrorInvalidValue;
  }
  if ((e = cudaFuncSetAttribute(corr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_corr)) != cudaSuccess) return e;
  const int nprob = a.U * a.nbins;
  const int iters = a.apply_only ? 0 : a.iterations;
  if (dual) {
    // chunk-major: K of a problem is formed once and serves every channel and every iteration (problems are independent, so
    // running the iterations of one chunk back to back is the same computation as estimate_Gn_'s iteration-major loop)
    for (int q0 = 0; q0 < nprob && iters > 0; q0 += chunk) {
      const int nq = (nprob - q0 < chunk) ? nprob - q0 : chunk;
      corr<<<nq, WPE_CORR_THREADS, sm_corr, st>>>(a, q0);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      (*launches)++;
      for (int it = 0; it < iters; it++) {
        k_wpe_resid<0><<<nq, 128, sm_resid, st>>>(a, q0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        chol<<<nq * C, chol_threads, sm_chol, st>>>(a, q0, C);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        (*launches) += 2;
      }
    }
  } else {
    for (int it = 0; it < iters; it++) {
      k_wpe_resid<0><<<nprob, 128, sm_resid, st>>>(a, 0);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      (*launches)++;
      for (int q0 = 0; q0 < nprob; q0 += chunk) {
        const int nq = (nprob - q0 < chunk) ? nprob - q0 : chunk;
        corr<<<dim3(8, nq), WPE_CORR_THREADS, sm_corr, st>>>(a, q0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        chol<<<nq * C, chol_threads, sm_chol, st>>>(a, q0, C);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        (*launches) += 2;
      }
    }
  }
  k_wpe_resid<1><<<a.G, 128, sm_resid, st>>>(a, 0);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  (*launches)++;
  return cudaSuccess;
}

cudaError_t launch_wpe(const WpeArgs& a, int chunk, int fp32, cudaStream_t st, int* launches) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  if (a.C < 1 || a.C > 8) return cudaErrorInvalidValue;
  return fp32 ? launch_wpe_t<float>(a, chunk, st, launches) : launch_wpe_t<double>(a, chunk, st, launches);
}

}  // namespace btkb
