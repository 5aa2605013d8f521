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

namespace btkb {

template <int M, int FB, int G>
__global__ void __launch_bounds__(G*(M / 8)) k_synthesis(SynthesisArgs a) {
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
  const int Tu = frames_of(len, D, a.laN, a.pdA);
  const int nbu = max(Tu - a.pdS, 0);
  const int t0 = tile * FB;
  const int tau0 = t0 + a.pdS - R * (m - 1) - (R - 1);  // first v frame of the tile (may be negative)

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
        if (va) ya = __ldg(a.Y + (size_t)ta * a.Gp + (size_t)u * K + k);
        if (vb) yb = __ldg(a.Y + (size_t)tb * a.Gp + (size_t)u * K + k);
        if (edge) { ya.y = 0.f; yb.y = 0.f; }
        if (cj) { ya.y = -ya.y; yb.y = -yb.y; }
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
    const int t = t0 + tl;
    if (t >= a.nb) break;
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
    a.out[(size_t)u * a.nb_stride + (size_t)t * D + (D - 1 - d)] = acc;
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

template <int M>
static cudaError_t launch_synthesis_m(const SynthesisArgs& a, cudaStream_t st) {
  using Plan = FftPlan<M>;
  constexpr int FB = (M >= 2048) ? 8 : 16;
  constexpr int G = (Plan::NT >= 128) ? 1 : (Plan::NT == 64 ? 2 : 4);
  const int R = 1 << a.r;
  const int NV = FB + R * (a.m - 1) + (R - 1);
  const int NVP = (NV + 1) & ~1;
  size_t smem = sizeof(float) * (size_t)NVP * M + sizeof(float2) * G * 2 * Plan::BUF + sizeof(float) * 32;
  auto kern = k_synthesis<M, FB, G>;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (a.nb <= 0) return cudaSuccess;
  dim3 grid((a.nb + FB - 1) / FB, a.U);
  kern<<<grid, G * Plan::NT, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_synthesis(const SynthesisArgs& a, cudaStream_t st) {
  switch (a.M) {
    case 256: return launch_synthesis_m<256>(a, st);
    case 512: return launch_synthesis_m<512>(a, st);
    case 1024: return launch_synthesis_m<1024>(a, st);
    case 2048: return launch_synthesis_m<2048>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace btkb
