// btkb_fused.cu — k_fused_analysis_nlms: OverSampledDFT analysis AND the per-bin GSC-NLMS recurrence in ONE kernel per frame-batch,
// the snapshots never leaving the SM (north star: "fused into one kernel per frame-batch that stages mic x bin tiles through shared
// memory").  Built to SETTLE the fusion question with a measurement; selected with BTKB_FUSED=1, off by default
// because it is slower than the two kernels it replaces (K1 + K4, profiles/r02i_fused.json).
//
// What it fuses (reference: one frame is pulled through the whole chain, modulated.cc:375-409 -> pybeamformer.py:659-734):
//   K1  k_analysis_r1<512, 4, ., ., PK>   polyphase fold + real-pair FFT + untangle + channel-0 energy      (btkb_analysis.cu)
//   K4  k_perbin<8, LMS, 0, PK>           Yc = v^H x, power-normalised leaky NLMS in projector form, output   (btkb_perbin.cu)
// One CTA = one utterance, persistent over its frames; 288 threads: warps 0..7 are four FFT groups (64 threads = one channel pair
// each, the K1 code), then threads 0..256 are the 257 bin chains (the K4 code) with their state in registers for the whole
// utterance.  Per iteration two frames: phase A (groups) writes the [2][8][257] snapshot tile to shared memory, phase B (chains)
// consumes it.  Samples are staged per super-tile of 8 frames with 16-byte cp.async copies (zero fill), like K1.
// Arithmetic and its order are those of K1 / K4 (same device functions, same frame pairing), so Y equals the unfused path BIT FOR
// BIT (tests/test_parity_gpu_r2.py::test_fused_analysis_nlms_equals_the_two_kernel_path).
// Restrictions: C = 8, M = 512, m = 4, r = 1, float32 samples, BTKB_BF_GSC_LMS, whole utterances.
#include "btkb_internal.h"
#include "btkb_fft.cuh"
#include "btkb_nlms_math.cuh"
#include "../../include/btkb.h"

namespace btkb {

namespace {
constexpr int FM = 512, FMT = 4, FC = 8, FNT = 64, FG = 4, FS8 = 8;            // M, taps factor, channels, threads per transform, groups, frames per super-tile
constexpr int FD = FM / 2, FK = FM / 2 + 1, FKP = 264;                         // frame shift, bins, padded bins per tile row
constexpr int FW = (FS8 - 1) * FD + FMT * FM;                                  // staged samples per channel and super-tile
constexpr int FTHREADS = 288;
__device__ __forceinline__ void bar_fft() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 threads of the four FFT groups
}  // namespace

__global__ void __launch_bounds__(FTHREADS, 1) k_fused_analysis_nlms(AnalysisArgs a, PerBinArgs b) {
  using Plan = FftPlan<FM>;
  constexpr int R0 = Plan::R0, NB = 8 / R0, SH = 4, NW = 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);                               // [8][FW] sample planes
  float2* fbuf = reinterpret_cast<float2*>(xs + (size_t)FC * FW);               // [FG][2][BUF]
  float2* Xt = fbuf + (size_t)FG * 2 * Plan::BUF;                               // [2][8][FKP] snapshot tile
  float* red = reinterpret_cast<float*>(Xt + 2 * FC * FKP);                     // [2 frames][NW] channel-0 energy partial sums
  const int u = blockIdx.x, tid = threadIdx.x;
  const bool fft_thread = tid < FG * FNT;
  const int grp = tid / FNT, tg = tid % FNT;                                    // (meaningful for fft threads)
  const int len = a.lengths[u];
  const int Tu = frames_of(len, FD, a.laN, b.pdA);
  // ---- analysis set-up (K1)
#define BTKB_SLOT(q) (((q) % NB) * R0 + (q) / NB)
  float hreg[8 * FMT];
  FftTwiddles<FM, +1> tw;
  float2* buf0 = fbuf + (grp * 2 + 0) * Plan::BUF;
  float2* buf1 = fbuf + (grp * 2 + 1) * Plan::BUF;
  if (fft_thread) {
#pragma unroll
    for (int q = 0; q < 8; q++)
#pragma unroll
      for (int k = 0; k < FMT; k++) hreg[BTKB_SLOT(q) * FMT + k] = __ldg(a.h + (tg + FNT * q) + k * FM);
    tw.init_from_table(tg, a.twtab);
  }
  const float* xa = a.x + ((size_t)u * FC + 2 * grp) * a.n_stride;
  const float* xb = xa + a.n_stride;
  const float* xsa = xs + (size_t)(2 * grp) * FW;
  const float* xsb = xsa + FW;
  // ---- per-bin set-up (K4): chain k = tid < 257
  const bool chain = tid < FK;
  const int k = chain ? tid : 0;
  const int g = u * FK + k;
  float2 w[FC], uw[FC];
#pragma unroll
  for (int c = 0; c < FC; c++) { w[c] = __ldg(b.W + (size_t)c * b.Gp + g); uw[c] = make_float2(0.f, 0.f); }
  float se = b.lms.init_diagonal_load, Eavg = b.lms.init_diagonal_load, gamma = b.lms.gamma;
  int n_updates = 0, slow_cnt = b.lms.slowdown_after + 1;
  const float one_m_beta = 1.0f - b.lms.beta, inv_sil = 1.0f / b.lms.sil_thresh;

  const int nst = (a.T + FS8 - 1) / FS8;
#pragma unroll 1
  for (int st = 0; st < nst; st++) {
    const int t0 = st * FS8;
    // ---- stage the super-tile's samples: all 288 threads, 16-byte zero-filling copies (previous super-tile fully consumed: last barrier)
    const long long w0 = a.w_base + (long long)t0 * FD - (long long)FMT * FM;
    for (int i = tid; i < FC * (FW / 4); i += FTHREADS) {
      const int c = i / (FW / 4), wq = 4 * (i - c * (FW / 4));
      const long long s = w0 + wq;
      const long long rem = (long long)len - s;
      const int nb = (s < 0 || rem <= 0) ? 0 : (rem >= 4 ? 16 : (int)rem * 4);
      const float* src = a.x + ((size_t)u * FC + c) * a.n_stride + (nb > 0 ? s : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(xs + (size_t)c * FW + wq)), "l"(src), "r"(nb) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < FS8 / 2; it++) {
      const int ta = t0 + 2 * it, tb = ta + 1;
      // ================= phase A: four groups, one channel pair each, frames (ta, tb) -> snapshot tile
      if (fft_thread) {
        const bool act0 = ta < a.T, act1 = tb < a.T;
        float2 v0[8], v1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { v0[i] = make_float2(0.f, 0.f); v1[i] = make_float2(0.f, 0.f); }
        if (act0) {
          const int base = (2 * it) * FD + FMT * FM - 1;
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int kk = 0; kk < FMT; kk++) {
              const float2 s = make_float2(xsa[base - (tg + FNT * q) - kk * FM], xsb[base - (tg + FNT * q) - kk * FM]);
              v0[BTKB_SLOT(q)] = f2_fma_s(s, hreg[BTKB_SLOT(q) * FMT + kk], v0[BTKB_SLOT(q)]);
              const int q1 = (q + SH >= 8) ? q + SH - 8 : q + SH;
              const int k1 = (q + SH >= 8) ? kk + 1 : kk;
              if (k1 < FMT) v1[BTKB_SLOT(q1)] = f2_fma_s(s, hreg[BTKB_SLOT(q1) * FMT + k1], v1[BTKB_SLOT(q1)]);
            }
#pragma unroll
          for (int q = 0; q < SH; q++) {
            const float2 s = make_float2(xsa[base + FD - (tg + FNT * q)], xsb[base + FD - (tg + FNT * q)]);
            v1[BTKB_SLOT(q)] = f2_fma_s(s, hreg[BTKB_SLOT(q) * FMT + 0], v1[BTKB_SLOT(q)]);
          }
        }
        fft_first_pass<FM, +1, true>(v0, buf0, tg);
        fft_first_pass<FM, +1, true>(v1, buf1, tg);
        bar_fft();
        auto sync = [] { bar_fft(); };
        FftPassChain<FM, +1, 0, decltype(sync), true>::run(v0, v1, buf0, buf1, tg, tw, sync);
        float e0 = 0.f, e1 = 0.f;
        const int ca = 2 * grp, cb = ca + 1;
#pragma unroll
        for (int q = 0; q <= 4; q++) {
          const int kb = tg + q * FNT;
          if (q == 4 && tg != 0) break;
          const float wgt = (kb == 0 || kb == FM / 2) ? 1.f : 2.f;
          const bool self = (kb == 0) || (q == 4);
          {
            const float2 zk = v0[q];
            const float2 zm = self ? zk : buf0[FM - kb];
            const float2 A = f2_scale(f2_add_conj(zk, zm), 0.5f);
            const float2 B = f2_scale_mi(f2_sub_conj(zk, zm), 0.5f);
            Xt[(0 * FC + ca) * FKP + kb] = A; Xt[(0 * FC + cb) * FKP + kb] = B;
            e0 = fmaf(wgt, fmaf(A.x, A.x, A.y * A.y), e0);
          }
          {
            const float2 zk = v1[q];
            const float2 zm = self ? zk : buf1[FM - kb];
            const float2 A = f2_scale(f2_add_conj(zk, zm), 0.5f);
            const float2 B = f2_scale_mi(f2_sub_conj(zk, zm), 0.5f);
            Xt[(1 * FC + ca) * FKP + kb] = A; Xt[(1 * FC + cb) * FKP + kb] = B;
            e1 = fmaf(wgt, fmaf(A.x, A.x, A.y * A.y), e1);
          }
        }
        if (grp == 0) {   // channel-0 energy: same reduction order as K1 (warp shuffle tree, then the two warp sums in order)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
          if ((tg & 31) == 0) { red[0 * NW + tg / 32] = e0; red[1 * NW + tg / 32] = e1; }
        }
      }
      __syncthreads();
      // ================= phase B: 257 chains, two frames of the NLMS recurrence (k_perbin<8, LMS, 0, PK = true>)
      if (chain) {
#pragma unroll 1
        for (int f = 0; f < 2; f++) {
          const int t = ta + f;
          if (t >= a.T) break;
          float2 x[FC];
#pragma unroll
          for (int c = 0; c < FC; c++) x[c] = Xt[(f * FC + c) * FKP + k];
          const float energy = (red[f * NW + 0] + red[f * NW + 1]) / (float)FM;
          float2 y = cdot<FC, true, true>(x, w);
          const bool live = t < Tu;
          if (--slow_cnt == 0) { gamma *= 0.5f; slow_cnt = b.lms.slowdown_after; }
          const bool adapt = energy > (Eavg * inv_sil);
          float2 nn = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < FC; c++) nn = f2_fma(x[c], x[c], nn);
          const float nx = nn.x + nn.y;
          float sub = (t > 0) ? fmaf(se, b.lms.beta, one_m_beta * nx) : nx;
          sub = fmaxf(sub, b.lms.energy_floor);
          if (adapt && live) {
            float n20, n21;
            nlms_adapt_step<FC, true>(x, w, uw, y, gamma, sub, b.lms.regularization_param, n20, n21);
            const float n2 = n20 + n21;
            if (n2 > b.lms.max_wa_l2norm) {
              const float cK = sqrtf(b.lms.max_wa_l2norm / n2);
#pragma unroll
              for (int c = 0; c < FC; c++) uw[c] = f2_scale(uw[c], cK);
            }
            se = sub;
            n_updates++;
          }
          if (t >= b.lms.min_frames) y = f2_sub(y, cdot<FC, false, true>(uw, x));
          Eavg = fmaf(Eavg, b.lms.beta, one_m_beta * energy);
          b.Y[(size_t)t * b.Gp + g] = live ? y : make_float2(0.f, 0.f);
        }
      }
      __syncthreads();   // the tile and the FFT buffers are free again
    }
  }
  if (chain) {
    if (b.UA != nullptr) {
#pragma unroll
      for (int c = 0; c < FC; c++) b.UA[(size_t)c * b.Gp + g] = uw[c];
    }
    if (k == 0 && b.stats_updates != nullptr) b.stats_updates[u] = (float)n_updates;
  }
#undef BTKB_SLOT
}

bool fused_supported(const AnalysisArgs& a, const PerBinArgs& b) {
  return a.M == FM && a.m == FMT && a.D == FD && a.C == FC && a.Crow == FC && a.x16 == nullptr && a.t_skip == 0 && b.kind == BTKB_BF_GSC_LMS && b.pf_kind == BTKB_PF_NONE &&
         b.ST == nullptr && b.t_base == 0;
}

cudaError_t launch_fused_analysis_nlms(const AnalysisArgs& a, const PerBinArgs& b, cudaStream_t st) {
  if (!fused_supported(a, b)) return cudaErrorInvalidValue;
  if (a.T <= 0 || a.U <= 0) return cudaSuccess;
  using Plan = FftPlan<FM>;
  const size_t smem = sizeof(float) * (size_t)FC * FW + sizeof(float2) * ((size_t)FG * 2 * Plan::BUF + 2 * FC * FKP) + sizeof(float) * 8;
  cudaError_t e = cudaFuncSetAttribute(k_fused_analysis_nlms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_fused_analysis_nlms<<<a.U, FTHREADS, smem, st>>>(a, b);
  return cudaGetLastError();
}

}  // namespace btkb
