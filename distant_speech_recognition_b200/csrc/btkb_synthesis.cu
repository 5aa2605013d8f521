// btkb_synthesis.cu — K5: batched OverSampledDFT polyphase synthesis (sm_100a).
//
// Replaces OverSampledDFTSynthesisBank::{update_buf_, next} (btk20_src/modulated/modulated.cc:551-612):
//   v_tau = Re(forward DFT(Y_tau))                      (e^{-2 pi i nk/M}; only the real part is kept, :563-564)
//   w_t[i] = sum_{k<m} g[(M-1-i) + kM] v_{t+pd-Rk}[i]   (frames with negative index are the zeroed buffer)
//   out_t[D-1-d] = sum_{s<R} w_{t-(R-1-s)}[d + sD]      (accumulated in float like the reference's gsl_vector_float)
// with the priming of pd frames (:573-578) folded into the index arithmetic.
//
// Y holds only the K = M/2+1 unique bins; the full spectrum is their Hermitian extension (the reference's beamformers
// fill it that way, beamformer.cc:1142-1149).  Because only Re(DFT) is kept, the imaginary parts of the DC and Nyquist
// bins cannot contribute and are dropped.  Two consecutive frames ride one complex transform:
//   DFT(Y_a + i Y_b) = v_a + i v_b   (v real).
//
// Mapping: one CTA = (utterance, tile of FB output blocks): FB + R(m-1) + (R-1) frames of v are produced into shared
// memory, then every output sample is 8 (= R m) MACs.
#include "btkb_internal.h"
#include "btkb_fft.cuh"
#include <cstdlib>

namespace btkb {

template <int M, int FB, int G>
__global__ void __launch_bounds__(G*(M / 8)) k_synthesis_generic(SynthesisArgs a) {
  using Plan = FftPlan<M>;
  constexpr int NT = Plan::NT;
  constexpr int R0 = Plan::R0;
  constexpr int NB = 8 / R0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tile = blockIdx.x, u = blockIdx.y;
  const int tid = threadIdx.x, grp = tid / NT, tg = tid % NT;
  const int R = 1 << a.r, m = a.m, D = a.D, K = a.K;
  const int NV = FB + R * (m - 1) + (R - 1);          // v frames needed by this tile
  const int NVP = (NV + 1) & ~1;                      // even (pairs)
  float* vs = reinterpret_cast<float*>(smem_raw);     // [NVP][M]
  float2* fbuf = reinterpret_cast<float2*>(vs + (size_t)NVP * M);
  float* red = reinterpret_cast<float*>(fbuf + G * 2 * Plan::BUF);
  float2* bufA = fbuf + (grp * 2 + 0) * Plan::BUF;
  float2* bufB = fbuf + (grp * 2 + 1) * Plan::BUF;

  const int len = a.lengths ? a.lengths[u] : a.n;
  const int Tu = a.tu ? a.tu[u] : frames_of(len, D, a.laN, a.pdA);   // absolute frame / block numbers below; t0 and the output rows are chunk-local
  const int nbu = max(Tu - a.pdS, 0);
  const int t0 = tile * FB - a.b_skip;   // chunk-local number of the tile's first block (-1: a discarded block that keeps the frame pairs aligned)
  const int tau0 = a.b_base + t0 + a.pdS - R * (m - 1) - (R - 1);  // first v frame of the tile (may be negative)

  FftTwiddles<M, -1> tw;
  tw.init(tg);

  for (int p0 = 0; p0 < NVP / 2; p0 += G) {
    const int pr = p0 + grp;
    const bool act = pr < NVP / 2;
    const int ta = tau0 + 2 * pr, tb = ta + 1;
    const bool va = act && ta >= 0 && ta < Tu, vb = act && tb >= 0 && tb < Tu;
    float2 v[8];
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
      for (int r = 0; r < R0; r++) {
        const int i = (tg + b * NT) + r * (M / R0);
        const int k = (i <= M / 2) ? i : M - i;
        const bool cj = i > M / 2;
        const bool edge = (k == 0) || (k == M / 2);
        float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
        if (va) ya = __ldg(a.Y + (size_t)(ta - a.y_base) * a.Gp + (size_t)u * K + k);
        if (vb) yb = __ldg(a.Y + (size_t)(tb - a.y_base) * a.Gp + (size_t)u * K + k);
        if (edge) { ya.y = 0.f; yb.y = 0.f; }
        if (cj) { ya.y = -ya.y; yb.y = -yb.y; }
        if (a.onesided > 0 && !edge) {  // Re IFFT of a spectrum whose upper half is zero = half the Hermitian one (DC, Nyquist whole)
          if (ta < a.onesided) { ya.x *= 0.5f; ya.y *= 0.5f; }
          if (tb < a.onesided) { yb.x *= 0.5f; yb.y *= 0.5f; }
        }
        v[b * R0 + r] = make_float2(ya.x - yb.y, ya.y + yb.x);  // ya + i yb
      }
    float2* Z = fft_run<M, -1>(v, bufA, bufB, tg, tw, [] { __syncthreads(); });
    if (act) {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int i = tg + q * NT;
        float2 z = Z[pidx(i)];
        vs[(size_t)(2 * pr) * M + i] = z.x;
        vs[(size_t)(2 * pr + 1) * M + i] = z.y;
      }
    }
    __syncthreads();
  }

  // ---- polyphase + overlap-add
  float sq = 0.f;
  const int nthreads = G * NT;
  for (int o = tid; o < FB * D; o += nthreads) {
    const int tl = o / D, d = o % D;
    if (t0 + tl >= a.nb) break;
    if (t0 + tl < 0) continue;
    const int t = a.b_base + t0 + tl;   // absolute block number
    float acc = 0.f;
    if (t < nbu) {
      for (int s = 0; s < R; s++) {
        const int i = d + s * D;
        // w_{t-(R-1-s)}[i]; frame index of v: t - (R-1-s) + pd - R k  -> tile slot = that - tau0
        const int tw_ = t - (R - 1 - s);
        float w = 0.f;
        if (tw_ >= 0) {
          for (int k = 0; k < m; k++) {
            const int slot = tw_ + a.pdS - R * k - tau0;
            w = fmaf(__ldg(a.g + (M - 1 - i) + k * M), vs[(size_t)slot * M + i], w);
          }
        }
        acc += w;
      }
      if (a.gain > 0) acc *= (float)a.gain;
    }
    a.out[(size_t)u * a.nb_stride + (size_t)(t0 + tl) * D + (D - 1 - d)] = acc;
    sq = fmaf(acc, acc, sq);
  }
  if (a.stats != nullptr) {
    // per-CTA sum of squares (double atomics: a statistic, not part of the parity-checked signal)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid & 31) == 0) red[tid / 32] = sq;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < (nthreads + 31) / 32; w++) s += (double)red[w];
      atomicAdd(a.stats + (size_t)u * 3, s);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast path (compile-time m = MT taps and R = RR = 2^r): the tile's Y rows are first staged in shared memory with
// asynchronous 8-byte copies (zero fill for frames outside the utterance), each group transforms TWO frame pairs per
// iteration (independent instruction streams) and writes the real sequences v back IN PLACE over the Y rows it consumed;
// the polyphase taps a thread needs (R m per output position) live in registers for all FB blocks of the tile.
//
// PK = true: packed 2 x fp32 arithmetic (btkb_f2.cuh) for the frame-pair combination Y_a + i Y_b, the transforms (the in-place pass
// chain of btkb_fft.cuh, whose registers hold the whole natural-order spectrum after the last pass) and the polyphase MACs (the
// R = 2 partial sums of an output sample are the two halves of one FFMA2 chain).  Same operations and roundings per component as
// PK = false.  The default since round 2 (0.143 vs 0.161 ms at configs[1] on B200); BTKB_SYNTHESIS_PACKED=0 selects the scalar kernel.
template <int M, int FB, int G, int MT, int RR, bool PK = false>
__global__ void __launch_bounds__(G*(M / 8)) k_synthesis_fast(SynthesisArgs a) {
  using Plan = FftPlan<M>;
  constexpr int NT = Plan::NT, R0 = Plan::R0, NB = 8 / R0, P = Plan::P;
  constexpr int D = M / RR, K = M / 2 + 1;
  constexpr int NV = FB + RR * (MT - 1) + (RR - 1);
  constexpr int NVP = (NV + 3) & ~3;                       // multiple of 4 frames (two pairs per group iteration)
  constexpr int ROW = ((K * 2 + 3) & ~3);                  // floats per staged row (>= 2K and >= M), 16 B multiple
  static_assert(ROW >= M, "row must hold M reals");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tile = blockIdx.x, u = blockIdx.y;
  const int tid = threadIdx.x, grp = tid / NT, tg = tid % NT;
  constexpr int NTHREADS = G * NT;
  float* rows = reinterpret_cast<float*>(smem_raw);        // [NVP][ROW]: Y (complex) on the way in, v (real) on the way out
  float2* fbuf = reinterpret_cast<float2*>(rows + (size_t)NVP * ROW);
  float* red = reinterpret_cast<float*>(fbuf + G * 2 * Plan::BUF);
  float2* buf0 = fbuf + (grp * 2 + 0) * Plan::BUF;
  float2* buf1 = fbuf + (grp * 2 + 1) * Plan::BUF;

  const int len = a.lengths[u];
  const int Tu = a.tu ? a.tu[u] : frames_of(len, D, a.laN, a.pdA);   // absolute frame / block numbers below; t0 and the output rows are chunk-local
  const int nbu = max(Tu - a.pdS, 0);
  const int t0 = tile * FB - a.b_skip;   // chunk-local number of the tile's first block (-1: a discarded block that keeps the frame pairs aligned)
  const int tau0 = a.b_base + t0 + a.pdS - RR * (MT - 1) - (RR - 1);

  // ---- stage Y rows
  for (int e = tid; e < NVP * K; e += NTHREADS) {
    const int fr = e / K, k = e - fr * K;
    const int tau = tau0 + fr;
    const bool ok = tau >= 0 && tau < Tu;   // (rows NV..NVP-1 are not used by this tile's blocks but are the pair partners of rows that are: every
                                            // frame must ride its transform with the same partner whatever the tiling, see SynthesisArgs::b_skip)
    const float2* src = ok ? a.Y + (size_t)(tau - a.y_base) * a.Gp + (size_t)u * K + k : a.Y;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(rows + (size_t)fr * ROW + 2 * k)), "l"(src), "r"(ok ? 8 : 0) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // polyphase taps for this thread's output positions d = tid + j NTHREADS
  constexpr int ND = (D + NTHREADS - 1) / NTHREADS;
  float gt[ND][RR][MT];
#pragma unroll
  for (int j = 0; j < ND; j++)
#pragma unroll
    for (int s2 = 0; s2 < RR; s2++)
#pragma unroll
      for (int k = 0; k < MT; k++) {
        const int d = tid + j * NTHREADS;
        gt[j][s2][k] = (d < D) ? __ldg(a.g + (M - 1 - (d + s2 * D)) + k * M) : 0.f;
      }
  FftTwiddles<M, -1> tw;
  tw.init_from_table(tg, a.twtab);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- transforms: group handles frame quads (2 pairs); v overwrites the Y rows of the same frames
  for (int q0 = 0; q0 < NVP / 4; q0 += G) {
    const int qd = q0 + grp;
    const bool act = qd < NVP / 4;
    float2 v0[8], v1[8];
    const float* ra = rows + (size_t)(4 * qd + 0) * ROW;
    const float* rb = rows + (size_t)(4 * qd + 1) * ROW;
    const float* rc = rows + (size_t)(4 * qd + 2) * ROW;
    const float* rd = rows + (size_t)(4 * qd + 3) * ROW;
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
      for (int r = 0; r < R0; r++) {
        const int i = (tg + b * NT) + r * (M / R0);
        const int k = (i <= M / 2) ? i : M - i;
        const bool cj = i > M / 2;
        const bool edge = (k == 0) || (k == M / 2);
        float2 ya = make_float2(0.f, 0.f), yb = ya, yc = ya, yd = ya;
        if (act) {
          ya = *reinterpret_cast<const float2*>(ra + 2 * k); yb = *reinterpret_cast<const float2*>(rb + 2 * k);
          yc = *reinterpret_cast<const float2*>(rc + 2 * k); yd = *reinterpret_cast<const float2*>(rd + 2 * k);
        }
        if (edge) { ya.y = 0.f; yb.y = 0.f; yc.y = 0.f; yd.y = 0.f; }
        if (cj) { ya.y = -ya.y; yb.y = -yb.y; yc.y = -yc.y; yd.y = -yd.y; }
        if (a.onesided > 0 && !edge) {  // see k_synthesis_generic
          const int tq = tau0 + 4 * qd;
          if (tq + 0 < a.onesided) { ya.x *= 0.5f; ya.y *= 0.5f; }
          if (tq + 1 < a.onesided) { yb.x *= 0.5f; yb.y *= 0.5f; }
          if (tq + 2 < a.onesided) { yc.x *= 0.5f; yc.y *= 0.5f; }
          if (tq + 3 < a.onesided) { yd.x *= 0.5f; yd.y *= 0.5f; }
        }
        if constexpr (PK) { v0[b * R0 + r] = f2_add_ib<+1>(ya, yb); v1[b * R0 + r] = f2_add_ib<+1>(yc, yd); }
        else {
        v0[b * R0 + r] = make_float2(ya.x - yb.y, ya.y + yb.x);
        v1[b * R0 + r] = make_float2(yc.x - yd.y, yc.y + yd.x);
        }
      }
    fft_first_pass<M, -1, PK>(v0, buf0, tg);
    fft_first_pass<M, -1, PK>(v1, buf1, tg);
    group_sync<NT, G>(grp);   // also: every thread of the group has consumed the Y rows of its quad (only this group reads and rewrites them)
    if constexpr (PK) {
      auto sync = [grp] { group_sync<NT, G>(grp); };
      FftPassChain<M, -1, 0, decltype(sync), PK>::run(v0, v1, buf0, buf1, tg, tw, sync);
      // v[r] = natural-order element tg + r M/8 of the two transforms (all eight still in registers): real parts to row a/c,
      // imaginary parts to row b/d
      if (act) {
        float* wa = rows + (size_t)(4 * qd + 0) * ROW; float* wb = rows + (size_t)(4 * qd + 1) * ROW;
        float* wc = rows + (size_t)(4 * qd + 2) * ROW; float* wd = rows + (size_t)(4 * qd + 3) * ROW;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int i = tg + r * (M / 8);
          wa[i] = v0[r].x; wb[i] = v0[r].y; wc[i] = v1[r].x; wd[i] = v1[r].y;
        }
      }
      group_sync<NT, G>(grp);   // the next quad's first pass reuses the group's buffers (the polyphase stage waits at the CTA barrier below)
    } else {
      int Ns = R0;
#pragma unroll
      for (int p = 0; p < P; p++) {
#pragma unroll
        for (int r = 0; r < 8; r++) { v0[r] = buf0[pidx(tg + r * (M / 8))]; v1[r] = buf1[pidx(tg + r * (M / 8))]; }
        group_sync<NT, G>(grp);
#pragma unroll
        for (int r = 1; r < 8; r++) { v0[r] = cmul(v0[r], tw.tw[p][r - 1]); v1[r] = cmul(v1[r], tw.tw[p][r - 1]); }
        dft8<-1>(v0);
        dft8<-1>(v1);
        if (p == P - 1) {
          // last pass: natural-order element tg + r M/8 -> real parts to row a/c, imaginary parts to row b/d
          if (act) {
            float* wa = rows + (size_t)(4 * qd + 0) * ROW; float* wb = rows + (size_t)(4 * qd + 1) * ROW;
            float* wc = rows + (size_t)(4 * qd + 2) * ROW; float* wd = rows + (size_t)(4 * qd + 3) * ROW;
#pragma unroll
            for (int r = 0; r < 8; r++) {
              const int i = tg + r * (M / 8);
              wa[i] = v0[r].x; wb[i] = v0[r].y; wc[i] = v1[r].x; wd[i] = v1[r].y;
            }
          }
        } else {
          stockham_store<8>(buf0, v0, tg, Ns);
          stockham_store<8>(buf1, v1, tg, Ns);
        }
        group_sync<NT, G>(grp);
        Ns *= 8;
      }
    }
  }
  __syncthreads();   // every group's v rows are in place: the polyphase stage reads all of them

  // ---- polyphase + overlap-add (taps in registers).  Output block tl of position d needs v rows tl + s2 + RR q (q < MT) at column
  // d + s2 D: consecutive blocks slide that window by ONE row, so the rows a position has read stay in registers (the block loop
  // is unrolled, the shift is register renaming) and each block costs RR new shared-memory reads instead of RR MT.  Arithmetic
  // and its order are those of the direct form: w_s2 = sum_k g[..] v[..] with k ascending, acc = (0 + w_0) + w_1.
  float sq = 0.f;
  constexpr int WL = RR * (MT - 1) + 1;
#pragma unroll
  for (int j = 0; j < ND; j++) {
    const int d = tid + j * NTHREADS;
    if (d < D) {
      float win[RR][WL];
#pragma unroll
      for (int s2 = 0; s2 < RR; s2++)
#pragma unroll
        for (int i = 0; i < WL - 1; i++) win[s2][i] = rows[(size_t)(s2 + i) * ROW + d + s2 * D];
#pragma unroll
      for (int tl = 0; tl < FB; tl++) {
        const int t = a.b_base + t0 + tl;   // absolute block number
#pragma unroll
        for (int s2 = 0; s2 < RR; s2++) win[s2][WL - 1] = rows[(size_t)(tl + s2 + WL - 1) * ROW + d + s2 * D];
        if (t0 + tl >= 0 && t0 + tl < a.nb) {
          float acc = 0.f;
          if (PK && RR == 2 && t < nbu) {
            // both partial sums w_{s2}, s2 = 0, 1, in one chain: taps (g[d + kM'], g[d + D + kM']) times (v_{t-1-2k}[d], v_{t-2k}[d + D]);
            // the s2 = 0 term does not exist for the first block of an utterance (twf = t - 1 < 0)
            float2 w2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < MT; k++)
              w2 = f2_fma(make_float2(gt[j][0][k], gt[j][1][k]), make_float2(win[0][RR * (MT - 1 - k)], win[RR - 1][RR * (MT - 1 - k)]), w2);
            acc = (t >= 1 ? w2.x : 0.f) + acc;   // acc = 0 + w_0, then + w_1: the scalar order
            acc += w2.y;
            if (a.gain > 0) acc *= (float)a.gain;
          } else if (t < nbu) {
#pragma unroll
            for (int s2 = 0; s2 < RR; s2++) {
              const int twf = t - (RR - 1 - s2);
              float w = 0.f;
              if (twf >= 0) {
#pragma unroll
                for (int k = 0; k < MT; k++) w = fmaf(gt[j][s2][k], win[s2][RR * (MT - 1 - k)], w);
              }
              acc += w;
            }
            if (a.gain > 0) acc *= (float)a.gain;
          }
          a.out[(size_t)u * a.nb_stride + (size_t)(t0 + tl) * D + (D - 1 - d)] = acc;
          sq = fmaf(acc, acc, sq);
        }
#pragma unroll
        for (int s2 = 0; s2 < RR; s2++)
#pragma unroll
          for (int i = 0; i < WL - 1; i++) win[s2][i] = win[s2][i + 1];
      }
    }
  }
  if (a.stats != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid & 31) == 0) red[tid / 32] = sq;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int w = 0; w < (NTHREADS + 31) / 32; w++) sum += (double)red[w];
      atomicAdd(a.stats + (size_t)u * 3, sum);
    }
  }
}

template <int M>
static cudaError_t launch_synthesis_fast(const SynthesisArgs& a, cudaStream_t st) {
  using Plan = FftPlan<M>;
  constexpr int FB = 16, MT = 4, RR = 2;
  constexpr int G = (Plan::NT >= 128) ? 1 : (Plan::NT == 64 ? 2 : 4);
  constexpr int NV = FB + RR * (MT - 1) + (RR - 1);
  constexpr int NVP = (NV + 3) & ~3;
  constexpr int ROW = (((M / 2 + 1) * 2 + 3) & ~3);
  size_t smem = sizeof(float) * (size_t)NVP * ROW + sizeof(float2) * G * 2 * Plan::BUF + sizeof(float) * 32;
  const bool packed = env_packed("BTKB_SYNTHESIS_PACKED");   // packed 2 x fp32 variant unless BTKB_SYNTHESIS_PACKED=0 (same results); read at every launch
  auto kern = packed ? k_synthesis_fast<M, FB, G, MT, RR, true> : k_synthesis_fast<M, FB, G, MT, RR, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (a.nb <= 0) return cudaSuccess;
  dim3 grid((a.nb + a.b_skip + FB - 1) / FB, a.U);
  kern<<<grid, G * Plan::NT, smem, st>>>(a);
  return cudaGetLastError();
}

template <int M>
static cudaError_t launch_synthesis_m(const SynthesisArgs& a, cudaStream_t st) {
  using Plan = FftPlan<M>;
  constexpr int FB = (M >= 2048) ? 8 : 16;
  constexpr int G = (Plan::NT >= 128) ? 1 : (Plan::NT == 64 ? 2 : 4);
  const int R = 1 << a.r;
  const int NV = FB + R * (a.m - 1) + (R - 1);
  const int NVP = (NV + 1) & ~1;
  size_t smem = sizeof(float) * (size_t)NVP * M + sizeof(float2) * G * 2 * Plan::BUF + sizeof(float) * 32;
  auto kern = k_synthesis_generic<M, FB, G>;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (a.nb <= 0) return cudaSuccess;
  dim3 grid((a.nb + a.b_skip + FB - 1) / FB, a.U);
  kern<<<grid, G * Plan::NT, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_synthesis(const SynthesisArgs& a, cudaStream_t st) {
  switch (a.M) {
    case 256: return (a.m == 4 && a.r == 1) ? launch_synthesis_fast<256>(a, st) : launch_synthesis_m<256>(a, st);
    case 512: return (a.m == 4 && a.r == 1) ? launch_synthesis_fast<512>(a, st) : launch_synthesis_m<512>(a, st);
    case 1024: return (a.m == 4 && a.r == 1) ? launch_synthesis_fast<1024>(a, st) : launch_synthesis_m<1024>(a, st);
    case 2048: return launch_synthesis_m<2048>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace btkb
