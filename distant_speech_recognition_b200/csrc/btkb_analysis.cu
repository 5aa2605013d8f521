// btkb_analysis.cu — K1: batched OverSampledDFT polyphase analysis (sm_100a).
//
// Replaces, for every (utterance, channel, frame) of a batch at once:
//   SampleFeature::next                        btk20_src/feature/feature.cc:605-649   (D-sample blocks, zero padding)
//   OverSampledDFTAnalysisBank::update_buffer_ btk20_src/modulated/modulated.cc:418-469 (look-ahead laN, zero tail)
//   OverSampledDFTAnalysisBank::next           modulated.cc:375-409  (polyphase fold + unnormalised backward DFT)
//   SnapShotArray::set_samples/update          btk20_src/beamformer/beamformer.cc:56-70 (channel -> bin-major transpose:
//                                              here a layout decision, not a copy)
//   MultiChannelSource.update_snapshot_array   btk20_src/lib/pybeamformer.py:263-277 (channel-0 frame energy)
//
// Closed form: frame t ends at sample n_t = (laN + t + 1) D - 1;
//   u[i] = sum_{k<m} h[i + kM] x[n_t - i - kM],  X_t[k] = sum_i u[i] e^{+2 pi i ik/M},  x = 0 outside the file.
//
// Mapping: one CTA = (utterance, channel PAIR, tile of FR frames).  The two real channels of a pair ride one complex
// transform (z = u_a + i u_b; X_a = (Z[k] + conj Z[M-k])/2, X_b = (Z[k] - conj Z[M-k])/(2i)), so C channels cost C/2
// complex FFTs.  The tile's samples ((FR-1) D + mM per channel, interleaved as float2 per pair) are staged once in
// shared memory; G transforms run concurrently (NT = M/8 threads each); the prototype taps a thread needs (8 m values)
// and its twiddles live in registers for the whole tile.
//
// HBM layout written: X[t][c][g], g = u K + k (complex64), row pitch Gp; E[t][u] = |x0^H x0| / M.
#include "btkb_internal.h"
#include "btkb_fft.cuh"

namespace btkb {

template <int M, int MT, int FR, int G>
__global__ void __launch_bounds__(G*(M / 8)) k_analysis(AnalysisArgs a) {
  using Plan = FftPlan<M>;
  constexpr int NT = Plan::NT;
  constexpr int R0 = Plan::R0;
  constexpr int NB = 8 / R0;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int tile = blockIdx.x, pair = blockIdx.y, u = blockIdx.z;
  const int tid = threadIdx.x;
  const int grp = tid / NT, tg = tid % NT;
  const int m = (MT > 0) ? MT : a.m;
  const int D = a.D;
  const int W = (FR - 1) * D + m * M;  // samples staged per channel
  float2* xs = reinterpret_cast<float2*>(smem_raw);                        // [W] (x_a, x_b)
  float2* fbuf = xs + W;                                                    // [G][2][BUF]
  float* red = reinterpret_cast<float*>(fbuf + G * 2 * Plan::BUF);          // [G][NT/32 or 1]
  float2* bufA = fbuf + (grp * 2 + 0) * Plan::BUF;
  float2* bufB = fbuf + (grp * 2 + 1) * Plan::BUF;

  const int t0 = tile * FR;
  const int ca = 2 * pair, cb = 2 * pair + 1;
  const bool has_b = cb < a.C;
  const int len = a.lengths ? a.lengths[u] : a.n;
  // first staged sample: n_{t0} - mM + 1
  const long long w0 = (long long)(a.laN + t0 + 1) * D - (long long)m * M;
  const float* xa = a.x + ((size_t)u * a.C + ca) * a.n_stride;
  const float* xb = a.x + ((size_t)u * a.C + (has_b ? cb : ca)) * a.n_stride;
  for (int w = tid; w < W; w += G * NT) {
    long long s = w0 + w;
    float va = 0.f, vb = 0.f;
    if (s >= 0 && s < len) { va = __ldg(xa + s); if (has_b) vb = __ldg(xb + s); }
    xs[w] = make_float2(va, vb);
  }

  // prototype taps this thread needs: element i = (tg + b NT) + r M/R0 sits in register slot b R0 + r
  float hreg[(MT > 0) ? 8 * MT : 1];
  if (MT > 0) {
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
      for (int r = 0; r < R0; r++) {
        int i = (tg + b * NT) + r * (M / R0);
#pragma unroll
        for (int k = 0; k < MT; k++) hreg[(b * R0 + r) * MT + k] = __ldg(a.h + i + k * M);
      }
  }
  FftTwiddles<M, +1> tw;
  tw.init(tg);
  __syncthreads();

  const int K = M / 2 + 1;
  for (int f0 = 0; f0 < FR; f0 += G) {
    const int f = f0 + grp;
    const int t = t0 + f;
    const bool active = (f < FR) && (t < a.T);
    // ---- polyphase fold straight into the leading-pass operand registers
    float2 v[8];
    {
      // sample x[n_t - i - kM] is at tile offset f D + mM - 1 - i - kM
      const int base = f * D + m * M - 1;
#pragma unroll
      for (int b = 0; b < NB; b++)
#pragma unroll
        for (int r = 0; r < R0; r++) {
          const int i = (tg + b * NT) + r * (M / R0);
          float2 acc = make_float2(0.f, 0.f);
          if (active) {
            if (MT > 0) {
#pragma unroll
              for (int k = 0; k < MT; k++) {
                float2 s = xs[base - i - k * M];
                float hv = hreg[(b * R0 + r) * MT + k];
                acc.x = fmaf(hv, s.x, acc.x); acc.y = fmaf(hv, s.y, acc.y);
              }
            } else {
              for (int k = 0; k < m; k++) {
                float2 s = xs[base - i - k * M];
                float hv = __ldg(a.h + i + k * M);
                acc.x = fmaf(hv, s.x, acc.x); acc.y = fmaf(hv, s.y, acc.y);
              }
            }
          }
          v[b * R0 + r] = acc;
        }
    }
    float2* Z = fft_run<M, +1>(v, bufA, bufB, tg, tw, [] { __syncthreads(); });
    // ---- untangle the two real channels, write snapshots, channel-0 energy
    float esum = 0.f;
    if (active) {
      const size_t rowa = ((size_t)t * a.C + ca) * a.Gp + (size_t)u * K;
      const size_t rowb = ((size_t)t * a.C + cb) * a.Gp + (size_t)u * K;
#pragma unroll
      for (int q = 0; q <= 4; q++) {
        int k = tg + q * NT;  // q < 4 covers 0..M/2-1 ; q == 4 only for k == M/2
        if (q == 4 && tg != 0) break;
        float2 zk = Z[pidx(k)];
        float2 zm = Z[pidx((M - k) & (M - 1))];
        float2 A = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        float2 d = make_float2(zk.x - zm.x, zk.y + zm.y);
        float2 B = make_float2(0.5f * d.y, -0.5f * d.x);
        if (a.gain > 0) { float gsc = (float)a.gain; A.x *= gsc; A.y *= gsc; B.x *= gsc; B.y *= gsc; }
        a.X[rowa + k] = A;
        if (has_b) a.X[rowb + k] = B;
        float wgt = (k == 0 || k == M / 2) ? 1.f : 2.f;
        esum = fmaf(wgt, fmaf(A.x, A.x, A.y * A.y), esum);
      }
    }
    if (pair == 0 && a.E != nullptr) {
      // fixed-order reduction over the NT threads of the group (deterministic)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
      constexpr int NW = (NT + 31) / 32;
      if ((tg & 31) == 0) red[grp * NW + tg / 32] = esum;
      __syncthreads();
      if (tg == 0 && active) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += red[grp * NW + w];
        a.E[(size_t)t * a.U + u] = s / (float)M;
      }
    }
    // the next iteration's first pass writes bufA: every thread has finished reading Z (bufA or bufB) only after this
    __syncthreads();
  }
}

template <int M, int MT>
static cudaError_t launch_analysis_m(const AnalysisArgs& a, cudaStream_t st) {
  using Plan = FftPlan<M>;
  constexpr int FR = 16;
  constexpr int G = (Plan::NT >= 128) ? 1 : (Plan::NT == 64 ? 2 : 4);
  const int m = (MT > 0) ? MT : a.m;
  size_t smem = sizeof(float2) * ((size_t)(FR - 1) * a.D + (size_t)m * M) + sizeof(float2) * G * 2 * Plan::BUF + sizeof(float) * G * 4;
  auto kern = k_analysis<M, MT, FR, G>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((a.T + FR - 1) / FR, (a.C + 1) / 2, a.U);
  kern<<<grid, G * Plan::NT, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_analysis(const AnalysisArgs& a, cudaStream_t st) {
#define BTKB_CASE(MM)                                                                  \
  case MM:                                                                             \
    return (a.m == 4) ? launch_analysis_m<MM, 4>(a, st) : launch_analysis_m<MM, 0>(a, st);
  switch (a.M) {
    BTKB_CASE(256) BTKB_CASE(512) BTKB_CASE(1024) BTKB_CASE(2048)
    default: return cudaErrorInvalidValue;
  }
#undef BTKB_CASE
}

}  // namespace btkb
