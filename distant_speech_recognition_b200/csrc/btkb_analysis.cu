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
#include <cstdlib>
#include <algorithm>

namespace btkb {

template <int M, int MT, int FR, int G>
__global__ void __launch_bounds__(G*(M / 8)) k_analysis_generic(AnalysisArgs a) {
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
  const long long w0 = a.w_base + (long long)t0 * D - (long long)m * M;
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
      const size_t rowa = ((size_t)t * a.Crow + ca) * a.Gp + (size_t)u * K;
      const size_t rowb = ((size_t)t * a.Crow + cb) * a.Gp + (size_t)u * K;
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


// ---------------------------------------------------------------------------------------------------------------
// Fast path (r = 1, i.e. D = M/2, compile-time tap count MT): each group of NT threads produces TWO consecutive frames
// per iteration.  Consecutive frames are D = 4 NT samples apart and a thread's polyphase indices are {tg + NT q}, so 28
// of the 32 samples frame t+1 needs are the ones the same thread already loaded for frame t: 36 shared-memory loads feed
// 64 (x2 channels) MACs.  The two transforms then advance pass by pass together (independent instruction streams, one
// in-place buffer each, every barrier covers two transforms).
//
// PK = true: the fold and the transforms use the packed 2 x fp32 instructions (btkb_f2.cuh).  The two real channels of a pair are
// the halves of one float2, so a tap MAC on both channels is one FFMA2 with the tap broadcast; results are bit-identical to
// PK = false.  The default since round 2 (0.499 vs 0.531 ms at configs[1] on B200); BTKB_ANALYSIS_PACKED=0 selects the scalar kernel.
// exact int16 -> fp32 of the two halves of a packed sample word (a | b << 16) without the conversion pipe: 0x4B000000 | (v ^ 0x8000) is the
// float 2^23 + (v + 32768); subtracting 2^23 + 32768 is exact
__device__ __forceinline__ float2 unpack_i16x2(uint32_t w) {
  const uint32_t t = w ^ 0x80008000u;
  const float lo = __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7610)) - 8421376.0f;
  const float hi = __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7632)) - 8421376.0f;
  return make_float2(lo, hi);
}

template <int M, int MT, int FR, int G, bool PK = false, bool I16 = false>
__global__ void __launch_bounds__(G*(M / 8), (FR <= 12 ? 4 : 3)) k_analysis_r1(AnalysisArgs a) {
  using Plan = FftPlan<M>;
  constexpr int NT = Plan::NT, R0 = Plan::R0, NB = 8 / R0, P = Plan::P;
  constexpr int D = M / 2, SH = 4;
  static_assert(FR % (2 * G) == 0, "tile must hold whole frame pairs per group");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pair = blockIdx.y, u = blockIdx.z;
  const int tid = threadIdx.x, grp = tid / NT, tg = tid % NT;
  constexpr int W = (FR - 1) * D + MT * M;
  // the two channels are staged as separate planes (16-byte cp.async chunks land without a register round trip)
  float* xsa = reinterpret_cast<float*>(smem_raw);                    // [W] channel a
  float* xsb = xsa + W;                                               // [W] channel b
  uint32_t* xs16 = reinterpret_cast<uint32_t*>(smem_raw);             // I16: [W] packed (a | b << 16)
  float2* fbuf = reinterpret_cast<float2*>(smem_raw + (I16 ? sizeof(uint32_t) * W : 2 * sizeof(float) * W));   // [G][2][BUF]
  float* red = reinterpret_cast<float*>(fbuf + G * 2 * Plan::BUF);    // [G][2][NW]
  float2* buf0 = fbuf + (grp * 2 + 0) * Plan::BUF;
  float2* buf1 = fbuf + (grp * 2 + 1) * Plan::BUF;

  const int ca = 2 * pair, cb = 2 * pair + 1;
  const bool has_b = cb < a.C;
  const int len = a.lengths[u];
  const float* xa = I16 ? nullptr : a.x + ((size_t)u * a.C + ca) * a.n_stride;
  const float* xb = I16 ? nullptr : a.x + ((size_t)u * a.C + (has_b ? cb : ca)) * a.n_stride;
  const int16_t* ya = I16 ? a.x16 + ((size_t)u * a.C + ca) * a.n16_stride : nullptr;
  const int16_t* yb = I16 ? a.x16 + ((size_t)u * a.C + (has_b ? cb : ca)) * a.n16_stride : nullptr;
  // slot(q): register slot of polyphase index i_q = tg + NT q  (q = b + r NB, slot = b R0 + r)
#define BTKB_SLOT(q) (((q) % NB) * R0 + (q) / NB)
  float hreg[8 * MT];
#pragma unroll
  for (int q = 0; q < 8; q++)
#pragma unroll
    for (int k = 0; k < MT; k++) hreg[BTKB_SLOT(q) * MT + k] = __ldg(a.h + (tg + NT * q) + k * M);
  FftTwiddles<M, +1> tw;
  tw.init_from_table(tg, a.twtab);
  constexpr int K = M / 2 + 1;
  constexpr int NW = (NT + 31) / 32;
  static_assert(W % 4 == 0, "tile length must be a multiple of 4 samples");

  // A CTA walks a.tiles_per_cta consecutive frame tiles of its (utterance, channel pair): prototype taps and twiddles are
  // set up once, the previous tile's last barrier protects the staged samples before they are overwritten.
  const int ntiles = (a.T + a.t_skip + FR - 1) / FR;
  const int tile_end = min((int)(blockIdx.x + 1) * a.tiles_per_cta, ntiles);
#pragma unroll 1
  for (int tile = blockIdx.x * a.tiles_per_cta; tile < tile_end; tile++) {
  const int t0 = tile * FR - a.t_skip;
  const long long w0 = a.w_base + (long long)t0 * D - (long long)MT * M;
  // Asynchronous 16-byte copies with zero fill (cp.async ... src-size): every thread puts ~W/(2 NT G) chunks per channel
  // in flight at once, samples outside [0, len) arrive as zeros (w0 and W are multiples of 4, rows are 16 B aligned).
  // The copies are committed in NIT groups, group i holding the samples iteration i needs beyond those of iteration i-1,
  // so the first frames start as soon as their window has landed while the rest of the tile is still in flight.
  constexpr int NIT = FR / (2 * G);
  if constexpr (I16) {
    // 16-bit PCM: 8 samples of each channel per thread and step (two 16-byte loads), interleaved into packed words with PRMT and stored as
    // two 16-byte shared-memory writes; samples outside [0, len) are zeros.  All loads of the tile are issued before the first store.
    static_assert(W % 8 == 0, "tile length must be a multiple of 8 samples");
    constexpr int NCH = (W / 8 + G * NT - 1) / (G * NT);
    uint4 ra[NCH], rb[NCH];
#pragma unroll
    for (int j = 0; j < NCH; j++) {
      const int w = 8 * (tid + j * G * NT);
      const long long s = w0 + w;
      ra[j] = make_uint4(0u, 0u, 0u, 0u); rb[j] = ra[j];
      if (w < W && s + 8 > 0 && s < (long long)len) {
        if (s >= 0 && s + 8 <= (long long)len) {
          ra[j] = __ldg(reinterpret_cast<const uint4*>(ya + s));
          if (has_b) rb[j] = __ldg(reinterpret_cast<const uint4*>(yb + s));
        } else {   // ragged edge of the utterance: sample by sample
          uint16_t ea[8], eb[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const long long si = s + i;
            const bool ok = si >= 0 && si < (long long)len;
            ea[i] = ok ? (uint16_t)ya[si] : (uint16_t)0; eb[i] = (ok && has_b) ? (uint16_t)yb[si] : (uint16_t)0;
          }
          ra[j] = make_uint4(ea[0] | (uint32_t)ea[1] << 16, ea[2] | (uint32_t)ea[3] << 16, ea[4] | (uint32_t)ea[5] << 16, ea[6] | (uint32_t)ea[7] << 16);
          rb[j] = make_uint4(eb[0] | (uint32_t)eb[1] << 16, eb[2] | (uint32_t)eb[3] << 16, eb[4] | (uint32_t)eb[5] << 16, eb[6] | (uint32_t)eb[7] << 16);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; j++) {
      const int w = 8 * (tid + j * G * NT);
      if (w < W) {
        uint4 lo, hi;   // word i = a_i | b_i << 16
        lo.x = __byte_perm(ra[j].x, rb[j].x, 0x5410); lo.y = __byte_perm(ra[j].x, rb[j].x, 0x7632);
        lo.z = __byte_perm(ra[j].y, rb[j].y, 0x5410); lo.w = __byte_perm(ra[j].y, rb[j].y, 0x7632);
        hi.x = __byte_perm(ra[j].z, rb[j].z, 0x5410); hi.y = __byte_perm(ra[j].z, rb[j].z, 0x7632);
        hi.z = __byte_perm(ra[j].w, rb[j].w, 0x5410); hi.w = __byte_perm(ra[j].w, rb[j].w, 0x7632);
        *reinterpret_cast<uint4*>(xs16 + w) = lo; *reinterpret_cast<uint4*>(xs16 + w + 4) = hi;
      }
    }
  } else {
#pragma unroll
  for (int gi = 0; gi < NIT; gi++) {
    const int lo = (gi == 0) ? 0 : (2 * G * gi - 1) * D + MT * M;
    const int hi = (2 * G * (gi + 1) - 1) * D + MT * M;
    for (int w = lo + 4 * tid; w < hi; w += 4 * G * NT) {
      const long long s = w0 + w;
      long long rem = (long long)len - s;                         // valid samples from s on
      int nb = (s < 0 || rem <= 0) ? 0 : (rem >= 4 ? 16 : (int)rem * 4);
      const float* pa = (nb > 0) ? xa + s : xa;
      const float* pb = (nb > 0) ? xb + s : xb;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(xsa + w)), "l"(pa), "r"(nb) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(xsb + w)), "l"(pb), "r"(has_b ? nb : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  }
#pragma unroll 1
  for (int it = 0; it < NIT; it++) {
    const int f0 = it * 2 * G;
    // groups 0..it must have landed: at most NIT-1-it of the most recent groups may still be pending
    if constexpr (!I16) switch (NIT - 1 - it) {
      case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
      case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
      case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
      case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
      case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
      case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
      case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
      default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
    __syncthreads();
    const int f = f0 + 2 * grp;
    const int ta = t0 + f, tb = ta + 1;
    const bool act0 = ta >= 0 && ta < a.T, act1 = tb < a.T;
    float2 v0[8], v1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { v0[i] = make_float2(0.f, 0.f); v1[i] = make_float2(0.f, 0.f); }
    if (ta < a.T && (PK || !(a.debug & 2))) {   // ta = -1 (t_skip): the discarded partner of local frame 0 is folded too; (the ablation hooks of BTKB_ANALYSIS_DEBUG exist in the scalar variant only)
      const int base = f * D + MT * M - 1;
#pragma unroll
      for (int q = 0; q < 8; q++)
#pragma unroll
        for (int k = 0; k < MT; k++) {
          const float2 s = I16 ? unpack_i16x2(xs16[base - (tg + NT * q) - k * M]) : make_float2(xsa[base - (tg + NT * q) - k * M], xsb[base - (tg + NT * q) - k * M]);
          const float h0 = hreg[BTKB_SLOT(q) * MT + k];
          if constexpr (PK) v0[BTKB_SLOT(q)] = f2_fma_s(s, h0, v0[BTKB_SLOT(q)]);
          else {
            v0[BTKB_SLOT(q)].x = fmaf(h0, s.x, v0[BTKB_SLOT(q)].x);
            v0[BTKB_SLOT(q)].y = fmaf(h0, s.y, v0[BTKB_SLOT(q)].y);
          }
          const int q1 = (q + SH >= 8) ? q + SH - 8 : q + SH;
          const int k1 = (q + SH >= 8) ? k + 1 : k;
          if (k1 < MT) {
            const float h1 = hreg[BTKB_SLOT(q1) * MT + k1];
            if constexpr (PK) v1[BTKB_SLOT(q1)] = f2_fma_s(s, h1, v1[BTKB_SLOT(q1)]);
            else {
              v1[BTKB_SLOT(q1)].x = fmaf(h1, s.x, v1[BTKB_SLOT(q1)].x);
              v1[BTKB_SLOT(q1)].y = fmaf(h1, s.y, v1[BTKB_SLOT(q1)].y);
            }
          }
        }
#pragma unroll
      for (int q = 0; q < SH; q++) {  // the D new samples of frame t+1 (tap block k = 0)
        const float2 s = I16 ? unpack_i16x2(xs16[base + D - (tg + NT * q)]) : make_float2(xsa[base + D - (tg + NT * q)], xsb[base + D - (tg + NT * q)]);
        const float h1 = hreg[BTKB_SLOT(q) * MT + 0];
        if constexpr (PK) v1[BTKB_SLOT(q)] = f2_fma_s(s, h1, v1[BTKB_SLOT(q)]);
        else {
          v1[BTKB_SLOT(q)].x = fmaf(h1, s.x, v1[BTKB_SLOT(q)].x);
          v1[BTKB_SLOT(q)].y = fmaf(h1, s.y, v1[BTKB_SLOT(q)].y);
        }
      }
    }
    // ---- two transforms, pass by pass, in place (one buffer each)
    fft_first_pass<M, +1, PK>(v0, buf0, tg);
    fft_first_pass<M, +1, PK>(v1, buf1, tg);
    __syncthreads();   // (per-group named barriers were measured here in round 2: 0.506 instead of 0.500 ms — the CTA barrier keeps the two groups' shared-memory bursts apart)
    if (PK || !(a.debug & 4)) {
      auto sync = [] { __syncthreads(); };
      FftPassChain<M, +1, 0, decltype(sync), PK>::run(v0, v1, buf0, buf1, tg, tw, sync);
    }
    // ---- untangle the channel pair, write snapshots, channel-0 energy.  After the last pass this thread holds
    // Z[tg + r NT] in v[r]; bins k = tg + q NT (q < 4) and k = M/2 (tg == 0, r = 4) are its own, only the partner
    // Z[M - k] (upper half, natural order in the buffer) comes from shared memory.
    float e0 = 0.f, e1 = 0.f;
#pragma unroll
    for (int q = 0; q <= 4; q++) {
      const int k = tg + q * NT;
      if (q == 4 && tg != 0) break;
      const float wgt = (k == 0 || k == M / 2) ? 1.f : 2.f;
      const bool self = (k == 0) || (q == 4);       // Z[M - k] is this thread's own value (k = 0 and k = M/2)
      if (act0) {
        const float2 zk = v0[q];
        const float2 zm = self ? zk : buf0[M - k];
        const float2 A = PK ? f2_scale(f2_add_conj(zk, zm), 0.5f) : make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 B = PK ? f2_scale_mi(f2_sub_conj(zk, zm), 0.5f) : make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        if (PK || !(a.debug & 1)) {
        a.X[((size_t)ta * a.Crow + ca) * a.Gp + (size_t)u * K + k] = A;
        if (has_b) a.X[((size_t)ta * a.Crow + cb) * a.Gp + (size_t)u * K + k] = B;
        }
        e0 = fmaf(wgt, fmaf(A.x, A.x, A.y * A.y), e0);
      }
      if (act1) {
        const float2 zk = v1[q];
        const float2 zm = self ? zk : buf1[M - k];
        const float2 A = PK ? f2_scale(f2_add_conj(zk, zm), 0.5f) : make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 B = PK ? f2_scale_mi(f2_sub_conj(zk, zm), 0.5f) : make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        if (PK || !(a.debug & 1)) {
        a.X[((size_t)tb * a.Crow + ca) * a.Gp + (size_t)u * K + k] = A;
        if (has_b) a.X[((size_t)tb * a.Crow + cb) * a.Gp + (size_t)u * K + k] = B;
        }
        e1 = fmaf(wgt, fmaf(A.x, A.x, A.y * A.y), e1);
      }
    }
    if (pair == 0 && a.E != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
      if ((tg & 31) == 0) { red[(grp * 2 + 0) * NW + tg / 32] = e0; red[(grp * 2 + 1) * NW + tg / 32] = e1; }
      __syncthreads();
      if (tg == 0) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) { s0 += red[(grp * 2 + 0) * NW + w]; s1 += red[(grp * 2 + 1) * NW + w]; }
        if (act0) a.E[(size_t)ta * a.U + u] = s0 / (float)M;
        if (act1) a.E[(size_t)tb * a.U + u] = s1 / (float)M;
      }
    }
    __syncthreads();  // untangle reads done before the next iteration's first pass overwrites the buffers
  }
  }  // tile loop
#undef BTKB_SLOT
}

static bool analysis_packed() { return env_packed("BTKB_ANALYSIS_PACKED"); }   // packed 2 x fp32 variant unless BTKB_ANALYSIS_PACKED=0

template <int M, int MT, int FR>
static cudaError_t launch_analysis_r1(const AnalysisArgs& a_in, cudaStream_t st) {
  AnalysisArgs a = a_in;
  using Plan = FftPlan<M>;
  constexpr int G = (Plan::NT >= 128) ? 1 : (Plan::NT == 64 ? 2 : 4);
  const bool i16 = a.x16 != nullptr;
  size_t smem = (i16 ? sizeof(uint32_t) : sizeof(float2)) * ((size_t)(FR - 1) * (M / 2) + (size_t)MT * M) + sizeof(float2) * G * 2 * Plan::BUF + sizeof(float) * G * 2 * ((Plan::NT + 31) / 32 + 1);   // red[G][2][NW] (round 2: was sized for NW <= 4, M = 2048 has 8 — found by compute-sanitizer)
  if (const char* e = getenv("BTKB_ANALYSIS_SMEM_PAD")) smem += (size_t)std::max(0, atoi(e));   // occupancy experiment: unused extra shared memory per CTA
  auto kern = i16 ? k_analysis_r1<M, MT, FR, G, true, true> : (analysis_packed() ? k_analysis_r1<M, MT, FR, G, true> : k_analysis_r1<M, MT, FR, G, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // tiles per CTA: amortise the per-CTA set-up while keeping >= ~8 waves of CTAs for balance
  const int ntiles = (a.T + a.t_skip + FR - 1) / FR;
  const long long ctas1 = (long long)ntiles * ((a.C + 1) / 2) * a.U;
  int tpc = 1;
  while (tpc < 8 && tpc * 2 <= ntiles && ctas1 / (tpc * 2) >= 148LL * 3 * 8) tpc *= 2;
  if (const char* e = getenv("BTKB_ANALYSIS_TPC")) { int v = atoi(e); if (v >= 1 && v <= 64) tpc = v; }
  a.tiles_per_cta = tpc;
  if (const char* e = getenv("BTKB_ANALYSIS_DEBUG")) a.debug = atoi(e);
  dim3 grid((ntiles + tpc - 1) / tpc, (a.C + 1) / 2, a.U);
  kern<<<grid, G * Plan::NT, smem, st>>>(a);
  return cudaGetLastError();
}

template <int M, int MT>
static cudaError_t launch_analysis_m(const AnalysisArgs& a, cudaStream_t st) {
  using Plan = FftPlan<M>;
  constexpr int FR = 16;
  constexpr int G = (Plan::NT >= 128) ? 1 : (Plan::NT == 64 ? 2 : 4);
  const int m = (MT > 0) ? MT : a.m;
  size_t smem = sizeof(float2) * ((size_t)(FR - 1) * a.D + (size_t)m * M) + sizeof(float2) * G * 2 * Plan::BUF + sizeof(float) * G * ((Plan::NT + 31) / 32 + 1);   // red[G][NW]
  auto kern = k_analysis_generic<M, MT, FR, G>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((a.T + FR - 1) / FR, (a.C + 1) / 2, a.U);
  kern<<<grid, G * Plan::NT, smem, st>>>(a);
  return cudaGetLastError();
}

static int analysis_tile_frames(bool i16) {  // tuning knob (frames per CTA tile): BTKB_ANALYSIS_FR=8|12|16; default 16 (3 CTAs/SM), 12 for 16-bit PCM
  static int v = -1;                       // input at M = 512 (its 24 KB sample tile lets 4 CTAs share an SM: 0.496 vs 0.511 ms)
  if (v < 0) { const char* e = getenv("BTKB_ANALYSIS_FR"); v = (e && atoi(e) == 8) ? 8 : (e && atoi(e) == 12) ? 12 : (e && atoi(e) == 16) ? 16 : 0; }
  return v ? v : (i16 ? 12 : 16);
}

cudaError_t launch_analysis(const AnalysisArgs& a, cudaStream_t st) {
#define BTKB_CASE(MM)                                                                  \
  case MM:                                                                             \
    if (a.m == 4 && a.D == MM / 2) return (analysis_tile_frames(a.x16 != nullptr) == 8) ? launch_analysis_r1<MM, 4, 8>(a, st) : (analysis_tile_frames(a.x16 != nullptr) == 12 && MM == 512) ? launch_analysis_r1<MM, 4, (MM == 512 ? 12 : 16)>(a, st) : launch_analysis_r1<MM, 4, 16>(a, st); \
    return (a.m == 4) ? launch_analysis_m<MM, 4>(a, st) : launch_analysis_m<MM, 0>(a, st);
  switch (a.M) {
    BTKB_CASE(256) BTKB_CASE(512) BTKB_CASE(1024) BTKB_CASE(2048)
    default: return cudaErrorInvalidValue;
  }
#undef BTKB_CASE
}

}  // namespace btkb
