// btkb_perbin.cu — K4 (+K2, K6): the fused per-bin spatial kernel (sm_100a).
//
// One thread owns one (utterance, bin) chain and walks the frames in order; a CTA owns TILE = 128 consecutive chains
// of the flattened g = u K + k axis.  Per frame the CTA's mic x bin tile ([C][128] complex64, C KiB) is brought from
// HBM into a 4-deep shared-memory ring by bulk asynchronous copies (cp.async.bulk + mbarrier complete_tx, i.e. the TMA
// engine; SASS: UBLKCP), issued by one elected thread, so the recurrences never wait on a global load and the loads
// are full 1 KiB row segments.  All per-chain state (weights, NLMS vector, sub-band energy, cross-spectral densities,
// covariance accumulators) lives in registers for the whole utterance.
//
// What it replaces (reference: btk20_src/):
//   static weights   SubbandDS::next            beamformer/beamformer.cc:1095-1157   y = wq^H x
//                    calc_gsc_output/SubbandGSC beamformer.cc:1208-1316              y = (wq - wl)^H x, DC bin: wq only
//                    SubbandMVDR[GSC]::next     beamformer.cc:2537-2587, 2719-2773   y = (wmvdr - wl)^H x, DC: wmvdr only
//   NLMS             SubbandGSCLMSBeamformer.__iter__  lib/pybeamformer.py:659-734   (power-normalised leaky LMS,
//                    silence gate, quadratic constraint, step halving, a-posteriori output)
//   post-filter      ZelinskiFilter_f / ZelinskiFilter / ZelinskiPostFilter::next  postfilter/postfilter.cc:8-43,57-219,424-491
//   covariance       SubbandSMIMVDRBeamformer.accu_stats_from_label  lib/pybeamformer.py:948-992  R += x x^H on noise frames
//
// NLMS in O(C) projector form (validated against the B-form oracle in tests/): with B the
// blocking matrix of v (B^T v = 0, B^H B = I), carry u = waH B^T (C-vector) instead of waH:
//   waH.Z = u.x ;  conj(Z) B^T = conj(x - C Yc v) ;  ||u~|| = ||wa~||.
// The (C-1) x C product per bin-frame disappears; waH = u conj(B) is recovered at export time (btkb_weights.cu).
#include "btkb_internal.h"
#include "btkb_fft.cuh"   // complex helpers
#include "../../include/btkb.h"

namespace btkb {

constexpr int TILE = 128;   // chains (threads) per CTA
constexpr int STAGES = 4;   // frames in flight per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// 1-D bulk asynchronous copy global -> shared, completion signalled on an mbarrier (bytes multiple of 16, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int C>
struct TileRing {
  float2* stage;     // [STAGES][C][TILE]
  uint64_t* full;    // [STAGES]
  const float2* X; int Gp, g0, T;
  __device__ __forceinline__ void init(unsigned char* smem, const float2* X_, int Gp_, int g0_, int T_) {
    stage = reinterpret_cast<float2*>(smem);
    full = reinterpret_cast<uint64_t*>(smem + sizeof(float2) * STAGES * C * TILE);
    X = X_; Gp = Gp_; g0 = g0_; T = T_;
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; s++) mbar_init(full + s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int t = 0; t < STAGES && t < T; t++) issue(t);
  }
  // called by ONE thread: fetch frame t into slot t % STAGES
  __device__ __forceinline__ void issue(int t) {
    const int s = t % STAGES;
    mbar_expect_tx(full + s, (uint32_t)(C * TILE * sizeof(float2)));
#pragma unroll
    for (int c = 0; c < C; c++)
      bulk_g2s(stage + ((size_t)s * C + c) * TILE, X + ((size_t)t * C + c) * Gp + g0, (uint32_t)(TILE * sizeof(float2)), full + s);
  }
  // all threads: wait for frame t, copy this thread's column to registers, release the slot, refill it
  __device__ __forceinline__ void fetch(int t, float2* x) {
    const int s = t % STAGES;
    mbar_wait(full + s, (uint32_t)((t / STAGES) & 1));
#pragma unroll
    for (int c = 0; c < C; c++) x[c] = stage[((size_t)s * C + c) * TILE + threadIdx.x];
    __syncthreads();  // every thread has its copy: the slot may be overwritten by the async proxy
    if (threadIdx.x == 0 && t + STAGES < T) issue(t + STAGES);
  }
};

constexpr int MODE_STATIC = 0, MODE_LMS = 1;

template <int C, int MODE, int PF>
__global__ void __launch_bounds__(TILE) k_perbin(PerBinArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  const int k = valid ? g - u * a.K : 0;
  const int len = a.lengths ? a.lengths[u] : 0;
  const int Tu = valid ? frames_of(len, a.D, a.laN, a.pdA) : 0;

  TileRing<C> ring;
  ring.init(smem_raw, a.X, a.Gp, g0, a.T);

  // ---- per-chain constants
  float2 w[C];     // static: effective weights (wq - wl, DC: wq); LMS: v = array manifold (wq)
  float2 ta[C];    // Zelinski time-alignment manifold (PF only; static modes)
#pragma unroll
  for (int c = 0; c < C; c++) {
    w[c] = __ldg(a.W + (size_t)c * a.Gp + g);
    if (PF) ta[c] = (a.TA != nullptr) ? __ldg(a.TA + (size_t)c * a.Gp + g) : w[c];
  }
  if (MODE == MODE_STATIC && a.WL != nullptr && k != 0) {
#pragma unroll
    for (int c = 0; c < C; c++) { float2 l = __ldg(a.WL + (size_t)c * a.Gp + g); w[c] = csub(w[c], l); }
  }

  // ---- NLMS state (pybeamformer.py:745-757 reset_stats)
  float2 uw[(MODE == MODE_LMS) ? C : 1];
  float se = a.lms.init_diagonal_load, Eavg = a.lms.init_diagonal_load, gamma = a.lms.gamma;
  int n_updates = 0;
  if (MODE == MODE_LMS) {
#pragma unroll
    for (int c = 0; c < C; c++) uw[c] = make_float2(0.f, 0.f);
  }
  // ---- Zelinski state: CSDs_ (beamformer.cc:874-887): upper triangle complex + real diagonal
  constexpr int NP = C * (C - 1) / 2;
  float2 csd[PF ? (NP > 0 ? NP : 1) : 1];
  float psd[PF ? C : 1];
  if (PF) {
#pragma unroll
    for (int i = 0; i < NP; i++) csd[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C; c++) psd[c] = 0.f;
  }

  float e_next = 0.f;
  if (MODE == MODE_LMS && a.T > 0) e_next = __ldg(a.E + u);

  for (int t = 0; t < a.T; t++) {
    float2 x[C];
    ring.fetch(t, x);
    float energy = e_next;
    if (MODE == MODE_LMS && t + 1 < a.T) e_next = __ldg(a.E + (size_t)(t + 1) * a.U + u);

    // upper branch: Yc = w^H x
    float2 y = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C; c++) { float2 p = cmulc(x[c], w[c]); y.x += p.x; y.y += p.y; }
    const bool live = t < Tu;

    if (MODE == MODE_LMS) {
      // pybeamformer.py:665-734 with isamp == t
      if (t > 0 && (t % a.lms.slowdown_after) == 0) gamma *= 0.5f;
      const bool adapt = energy > (Eavg / a.lms.sil_thresh);
      float nx = 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) nx = fmaf(x[c].x, x[c].x, fmaf(x[c].y, x[c].y, nx));
      float sub = (t > 0) ? fmaf(se, a.lms.beta, (1.0f - a.lms.beta) * nx) : nx;
      sub = fmaxf(sub, a.lms.energy_floor);
      if (adapt && live) {
        float2 ux = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C; c++) { float2 p = cmul(uw[c], x[c]); ux.x += p.x; ux.y += p.y; }
        const float2 epa = csub(y, ux);
        const float alphaK = gamma / sub;
        const float2 cy = make_float2((float)C * y.x, (float)C * y.y);
        const float2 ea = make_float2(epa.x * alphaK, epa.y * alphaK);
        const float leak = (a.lms.regularization_param > 0.f) ? alphaK * a.lms.regularization_param : 0.f;
        float n2 = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
          float2 q = csub(x[c], cmul(cy, w[c]));      // (Q x)_c = x_c - C Yc v_c
          float2 d = cmulc(ea, q);                    // e a conj(q)
          float2 un = make_float2(uw[c].x + d.x - leak * uw[c].x, uw[c].y + d.y - leak * uw[c].y);
          uw[c] = un;
          n2 = fmaf(un.x, un.x, fmaf(un.y, un.y, n2));
        }
        if (n2 > a.lms.max_wa_l2norm) {
          const float cK = sqrtf(a.lms.max_wa_l2norm / n2);
#pragma unroll
          for (int c = 0; c < C; c++) { uw[c].x *= cK; uw[c].y *= cK; }
        }
        se = sub;
        n_updates++;
      }
      if (t >= a.lms.min_frames) {
        float2 ux = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C; c++) { float2 p = cmul(uw[c], x[c]); ux.x += p.x; ux.y += p.y; }
        y = csub(y, ux);
      }
      Eavg = fmaf(Eavg, a.lms.beta, (1.0f - a.lms.beta) * energy);
    }

    if (PF) {
      // ZelinskiFilter_f (postfilter.cc:57-140); alpha = 0 for the first two frames (postfilter.cc:460-463)
      const float al = (t >= 2) ? a.pf_alpha : 0.f;
      float2 z[C];
#pragma unroll
      for (int c = 0; c < C; c++) z[c] = cmulc(x[c], ta[c]);
      float2 s = make_float2(0.f, 0.f);
      float den = 0.f;
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C - 1; i++)
#pragma unroll
        for (int j = i + 1; j < C; j++) {
          float2 zz = cmulc(z[i], z[j]);
          float2 ph = (al > 0.f) ? make_float2(fmaf(al, csd[idx].x, (1.f - al) * zz.x), fmaf(al, csd[idx].y, (1.f - al) * zz.y)) : zz;
          if (live) csd[idx] = ph;
          s.x += ph.x; s.y += ph.y;
          idx++;
        }
#pragma unroll
      for (int c = 0; c < C; c++) {
        float pz = fmaf(z[c].x, z[c].x, z[c].y * z[c].y);
        float ps = (al > 0.f) ? fmaf(al, psd[c], (1.f - al) * pz) : pz;
        if (live) psd[c] = ps;
        den += ps;
      }
      float num = (a.pf_type & 1) ? fmaxf(s.x, 0.f) : sqrtf(fmaf(s.x, s.x, s.y * s.y));
      float Wf = (num / den) * (2.0f / ((float)C - 1.0f));
      Wf = (Wf >= 1.0f) ? 1.0f : Wf;
      Wf = (Wf < 1.0e-4f) ? 1.0e-4f : Wf;
      if (!(den > 0.f)) Wf = 1.0e-4f;  // all-zero snapshot: the reference computes 0/0 = NaN -> fails both tests; keep finite
      if (t - 1 >= a.pf_min_frames) { y.x *= Wf; y.y *= Wf; }
      if (a.PFW != nullptr && valid) a.PFW[(size_t)t * a.Gp + g] = Wf;
    }

    if (valid) a.Y[(size_t)t * a.Gp + g] = live ? y : make_float2(0.f, 0.f);
  }

  if (MODE == MODE_LMS && valid) {
    if (a.UA != nullptr) {
#pragma unroll
      for (int c = 0; c < C; c++) a.UA[(size_t)c * a.Gp + g] = uw[c];
    }
    if (k == 0 && a.stats_updates != nullptr) a.stats_updates[u] = (float)n_updates;
  }
}

// K2: R[u][k] += x x^H over the frames flagged in noise_mask (pybeamformer.py:976-982)
template <int C>
__global__ void __launch_bounds__(TILE) k_covariance(PerBinArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  TileRing<C> ring;
  ring.init(smem_raw, a.X, a.Gp, g0, a.T);
  constexpr int NP = C * (C - 1) / 2;
  float2 off[NP > 0 ? NP : 1];
  float dg[C];
#pragma unroll
  for (int i = 0; i < NP; i++) off[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < C; c++) dg[c] = 0.f;
  unsigned char m_next = (a.T > 0) ? a.noise_mask[u] : 0;
  for (int t = 0; t < a.T; t++) {
    float2 x[C];
    ring.fetch(t, x);
    const unsigned char mk = m_next;
    if (t + 1 < a.T) m_next = a.noise_mask[(size_t)(t + 1) * a.U + u];
    if (mk) {
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C; i++) {
        dg[i] = fmaf(x[i].x, x[i].x, fmaf(x[i].y, x[i].y, dg[i]));
#pragma unroll
        for (int j = i + 1; j < C; j++) { float2 p = cmulc(x[i], x[j]); off[idx].x += p.x; off[idx].y += p.y; idx++; }
      }
    }
  }
  if (valid) {
    int idx = 0;
#pragma unroll
    for (int i = 0; i < C; i++) {
      a.R[(size_t)(i * C + i) * a.Gp + g] = make_float2(dg[i], 0.f);
#pragma unroll
      for (int j = i + 1; j < C; j++) {
        a.R[(size_t)(i * C + j) * a.Gp + g] = off[idx];
        a.R[(size_t)(j * C + i) * a.Gp + g] = make_float2(off[idx].x, -off[idx].y);
        idx++;
      }
    }
  }
}

template <int C>
static size_t ring_smem() { return sizeof(float2) * STAGES * C * TILE + sizeof(uint64_t) * STAGES + 64; }

template <int C>
static cudaError_t launch_perbin_c(const PerBinArgs& a, cudaStream_t st) {
  const size_t smem = ring_smem<C>();
  const int grid = (a.G + TILE - 1) / TILE;
  const bool lms = a.kind == BTKB_BF_GSC_LMS;
  const bool pf = a.pf_kind == BTKB_PF_ZELINSKI;
#define BTKB_LAUNCH(MODE_, PF_)                                                                  \
  do {                                                                                           \
    auto kern = k_perbin<C, MODE_, PF_>;                                                         \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                              \
    kern<<<grid, TILE, smem, st>>>(a);                                                           \
    return cudaGetLastError();                                                                   \
  } while (0)
  if (lms) { if (pf) return cudaErrorInvalidValue; BTKB_LAUNCH(MODE_LMS, 0); }
  if (pf) BTKB_LAUNCH(MODE_STATIC, 1);
  BTKB_LAUNCH(MODE_STATIC, 0);
#undef BTKB_LAUNCH
}

cudaError_t launch_perbin(const PerBinArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  switch (a.C) {
    case 2: return launch_perbin_c<2>(a, st);
    case 4: return launch_perbin_c<4>(a, st);
    case 8: return launch_perbin_c<8>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int C>
static cudaError_t launch_cov_c(const PerBinArgs& a, cudaStream_t st) {
  const size_t smem = ring_smem<C>();
  auto kern = k_covariance<C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(a.G + TILE - 1) / TILE, TILE, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_covariance(const PerBinArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  switch (a.C) {
    case 2: return launch_cov_c<2>(a, st);
    case 4: return launch_cov_c<4>(a, st);
    case 8: return launch_cov_c<8>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace btkb
