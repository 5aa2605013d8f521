// btkb_perbin.cu — K4 (+K2, K6): the fused per-bin spatial kernel (sm_100a).
//
// One thread owns one (utterance, bin) chain and walks the frames in order; a CTA owns TILE = 64 consecutive chains
// of the flattened g = u K + k axis.  The CTA's mic x bin tile of FCH = 2 frames ([2][C][64] complex64, 8 KiB at C = 8) is
// brought from HBM into a 3-slot shared-memory ring by ONE tensor-map TMA instruction per slot
// (cp.async.bulk.tensor.2d + mbarrier complete_tx; SASS: UTMALDG.2D), issued by one elected thread; slots are released
// through per-slot `empty` mbarriers (one arrival per warp), so the frame loop has no CTA-wide barrier and the
// recurrences never wait on a global load.  All per-chain state (weights, NLMS vector, sub-band energy, cross-spectral
// densities, covariance accumulators) lives in registers for the whole utterance.
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
#include "btkb_tile_ring.cuh"
#include "btkb_fft.cuh"   // complex helpers
#include "btkb_nlms_math.cuh"
#include "../../include/btkb.h"
#include <cstdlib>

// Channel counts: every C in 1..8 is instantiated (the reference takes any number of set_channel() calls, beamformer.cc:1017-1021).
// To keep the build parallel this file is compiled twice: BTKB_CSET=0 -> C in {2, 4, 8} and the public launch_* entry points,
// BTKB_CSET=1 -> C in {1, 3, 5, 6, 7} behind launch_*_cset1 (csrc/Makefile).
#ifndef BTKB_CSET
#define BTKB_CSET 0
#endif

namespace btkb {

cudaError_t launch_perbin_cset1(const PerBinArgs& a, cudaStream_t st);
cudaError_t launch_covariance_cset1(const PerBinArgs& a, cudaStream_t st);
cudaError_t launch_spectral_recursion_cset1(const PerBinArgs& a, float mu, int noconj, cudaStream_t st);

constexpr int MODE_STATIC = 0, MODE_LMS = 1;

// PK = true (the default; BTKB_PERBIN_PACKED=0 selects the scalar form): the complex arithmetic of the frame loop in packed 2 x fp32
// instructions — 2 FFMA2 per complex MAC instead of 4 FFMA, about half the fp32 issue slots of the recurrence.  Bit-identical
// results (tests/test_fft_packed_host.py checks every packed formula against the scalar expression it replaces).
template <int C, int MODE, int PF, bool PK = false>
__global__ void __launch_bounds__(TILE) k_perbin(const __grid_constant__ CUtensorMap tmX, PerBinArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  const int k = valid ? g - u * a.K : 0;
  const int len = a.lengths ? a.lengths[u] : 0;
  const int Tu = valid ? (a.tu ? a.tu[u] - a.t_base : frames_of(len, a.D, a.laN, a.pdA)) : 0;   // live frames of this chain among the a.T local ones
  // carried state of a streamed chunk (btkb_stream_submit): row i of ST belongs to this chain at ST[i Gp + g]
  const bool has_st = a.ST != nullptr && valid;
  float* const stp = a.ST + (valid ? g : 0);   // (never formed from a null test: ptxas speculated the state loads past a pointer-valued select)
  const bool st_load = has_st && a.st_load != 0;

  TileRing<C> ring;
  ring.init(smem_raw, &tmX, g0, a.T);

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
  if (MODE == MODE_STATIC && a.normalize_weight && k != 0 && a.kind != BTKB_BF_DS) {
    // calc_gsc_output(normalizeWeight = true): w <- w / (||w|| C) (beamformer.cc:1230-1236); the DC bin bypasses it
    float nrm = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) nrm = fmaf(w[c].x, w[c].x, fmaf(w[c].y, w[c].y, nrm));
    const float sc = 1.0f / (sqrtf(nrm) * (float)C);
#pragma unroll
    for (int c = 0; c < C; c++) { w[c].x *= sc; w[c].y *= sc; }
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
  float2 csd[(PF == 1) ? (NP > 0 ? NP : 1) : 1];
  float psd[PF ? C : 1];
  if (PF) {
    if (PF == 1) {
#pragma unroll
      for (int i = 0; i < NP; i++) csd[i] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < C; c++) psd[c] = 0.f;
  }
  // ---- McCowan / Lefkimmiatis state.  The per-pair CSD recursion phi_ij <- al phi_ij + (1-al) z_i conj(z_j) is linear, so
  // the two pair sums the filters need, S' = sum q'_ij phi_ij and S'' = sum q''_ij phi_ij, obey the same recursion and are
  // carried as two complex scalars; the bin's constants q', rho', q'', rho'' (btkb_postfilter.cu) sit in this thread's
  // column of shared memory, [entry][TILE], read conflict-free every frame.
  constexpr int NQ = (PF == 3) ? 2 * (NP + C) : (NP + C);
  float2* pfc = reinterpret_cast<float2*>(smem_raw + ring_bytes<C>()) + threadIdx.x;
  float2 S1 = make_float2(0.f, 0.f), S2 = make_float2(0.f, 0.f);
  float lam_inv = 1.0f;
  if (PF >= 2) {
    for (int q = 0; q < NQ; q++) pfc[q * TILE] = __ldg(a.PFQ + (size_t)q * a.K + k);
    if (PF == 3) lam_inv = (k < a.pf_fbin1) ? 1.0f : 1.0f / __ldg(a.LAM + g);
  }

  float e_next = 0.f;
  if (MODE == MODE_LMS && a.T > 0) e_next = __ldg(a.E + u);
  int slow_cnt = a.lms.slowdown_after + 1;  // first halving at t == slowdown_after
  const float one_m_beta = 1.0f - a.lms.beta;
  const float inv_sil = 1.0f / a.lms.sil_thresh;
  if (st_load) {   // resume where the previous chunk of the stream stopped
    const size_t gp = (size_t)a.Gp;
    if (MODE == MODE_LMS) {
      se = stp[0 * gp]; Eavg = stp[1 * gp]; gamma = stp[2 * gp]; slow_cnt = __float_as_int(stp[3 * gp]); n_updates = __float_as_int(stp[4 * gp]);
#pragma unroll
      for (int c = 0; c < C; c++) uw[c] = a.UA[(size_t)c * a.Gp + g];
    }
    if (PF == 1) {
#pragma unroll
      for (int i = 0; i < NP; i++) csd[i] = make_float2(stp[(8 + 2 * i) * gp], stp[(9 + 2 * i) * gp]);
    }
    if (PF >= 2) { S1 = make_float2(stp[8 * gp], stp[9 * gp]); S2 = make_float2(stp[10 * gp], stp[11 * gp]); }
    if (PF) {
#pragma unroll
      for (int c = 0; c < C; c++) psd[c] = stp[((PF == 1 ? 8 + 2 * NP : 12) + c) * gp];
    }
  }

  for (int tl = 0; tl < a.T; tl++) {
    const int t = a.t_base + tl;   // absolute frame number: what the recurrences test
    float2 x[C];
    ring.fetch(tl, a.T, x);
    float energy = e_next;
    if (MODE == MODE_LMS && tl + 1 < a.T) e_next = __ldg(a.E + (size_t)(tl + 1) * a.U + u);

    // upper branch: Yc = w^H x
    float2 y = cdot<C, true, PK>(x, w);
    const bool live = tl < Tu;

    if (MODE == MODE_LMS) {
      // pybeamformer.py:665-734 with isamp == t
      if (--slow_cnt == 0) { gamma *= 0.5f; slow_cnt = a.lms.slowdown_after; }  // isamp > 0 and isamp % slowdown_after == 0
      const bool adapt = energy > (Eavg * inv_sil);
      float nx0 = 0.f, nx1 = 0.f;
      if constexpr (PK) {
        float2 nn = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C; c++) nn = f2_fma(x[c], x[c], nn);
        nx0 = nn.x; nx1 = nn.y;
      } else {
#pragma unroll
      for (int c = 0; c < C; c++) { nx0 = fmaf(x[c].x, x[c].x, nx0); nx1 = fmaf(x[c].y, x[c].y, nx1); }
      }
      const float nx = nx0 + nx1;
      float sub = (t > 0) ? fmaf(se, a.lms.beta, one_m_beta * nx) : nx;
      sub = fmaxf(sub, a.lms.energy_floor);
      if (adapt && live) {
        float n20, n21;
        nlms_adapt_step<C, PK>(x, w, uw, y, gamma, sub, a.lms.regularization_param, n20, n21);
        const float n2 = n20 + n21;
        if (n2 > a.lms.max_wa_l2norm) {
          const float cK = sqrtf(a.lms.max_wa_l2norm / n2);
#pragma unroll
          for (int c = 0; c < C; c++) { if constexpr (PK) uw[c] = f2_scale(uw[c], cK); else { uw[c].x *= cK; uw[c].y *= cK; } }
        }
        se = sub;
        n_updates++;
      }
      if (t >= a.lms.min_frames) {
        const float2 ux = cdot<C, false, PK>(uw, x);
        y = PK ? f2_sub(y, ux) : csub(y, ux);
      }
      Eavg = fmaf(Eavg, a.lms.beta, one_m_beta * energy);
    }

    if constexpr (PF >= 2) {
      // McCowanPostFilter::post_filtering_ (postfilter.cc:826-885) / LefkimmiatisPostFilter::post_filtering_ (:1090-1160);
      // alpha = 0 while frame_no_ <= 0, i.e. for the first two frames, as in Zelinski's filter
      const float al = (t >= 2) ? a.pf_alpha : 0.f;
      const float be = 1.0f - al;
      float2 z[C];
#pragma unroll
      for (int c = 0; c < C; c++) z[c] = PK ? f2_cmulc(x[c], ta[c]) : cmulc(x[c], ta[c]);
      float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C - 1; i++)
#pragma unroll
        for (int j = i + 1; j < C; j++) {
          if constexpr (PK) {
            const float2 zz = f2_cmulc(z[i], z[j]);
            a1 = f2_cmac(a1, zz, pfc[idx * TILE]);
            if (PF == 3) a2 = f2_cmac(a2, zz, pfc[(NP + C + idx) * TILE]);
          } else {
          const float2 zz = cmulc(z[i], z[j]);
          cmac(a1, zz, pfc[idx * TILE]);
          if (PF == 3) cmac(a2, zz, pfc[(NP + C + idx) * TILE]);
          }
          idx++;
        }
      const float2 s1 = (al > 0.f) ? make_float2(fmaf(al, S1.x, be * a1.x), fmaf(al, S1.y, be * a1.y)) : a1;
      const float2 s2 = (al > 0.f) ? make_float2(fmaf(al, S2.x, be * a2.x), fmaf(al, S2.y, be * a2.y)) : a2;
      float2 r1 = make_float2(0.f, 0.f), r2 = make_float2(0.f, 0.f);
      float den = 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const float pz = fmaf(z[c].x, z[c].x, z[c].y * z[c].y);
        const float ps = (al > 0.f) ? fmaf(al, psd[c], be * pz) : pz;
        if (live) psd[c] = ps;
        den += ps;
        const float2 h1 = pfc[(NP + c) * TILE];
        r1.x = fmaf(ps, h1.x, r1.x); r1.y = fmaf(ps, h1.y, r1.y);
        if (PF == 3) { const float2 h2 = pfc[(2 * NP + C + c) * TILE]; r2.x = fmaf(ps, h2.x, r2.x); r2.y = fmaf(ps, h2.y, r2.y); }
      }
      if (live) { S1 = s1; S2 = s2; }
      const float pn = 2.0f / ((float)C * ((float)C - 1.0f));
      const float2 cs = make_float2(s1.x - r1.x, s1.y - r1.y);  // sum over pairs of (phi_ij - R' (phi_ii + phi_jj)/2) / (1 - R')
      const float phi_ss = pn * ((a.pf_type & 1) ? cs.x : sqrtf(fmaf(cs.x, cs.x, cs.y * cs.y)));
      float Wf;
      if (PF == 2) {
        Wf = phi_ss / (den / (float)C);
      } else {
        const float2 cv = make_float2(r2.x - s2.x, r2.y - s2.y);  // sum over pairs of ((phi_ii + phi_jj)/2 - phi_ij) / (1 - R'')
        const float phi_vv = pn * ((a.pf_type & 1) ? cv.x : sqrtf(fmaf(cv.x, cv.x, cv.y * cv.y)));
        Wf = phi_ss / fmaf(phi_vv, lam_inv, phi_ss);
      }
      Wf = (Wf > 1.0f) ? 1.0f : Wf;
      Wf = (Wf < 1.0e-4f) ? 1.0e-4f : Wf;
      if (!(Wf == Wf)) Wf = 1.0e-4f;  // all-zero snapshot: 0/0 in the reference; keep the output finite
      if (t - 1 >= a.pf_min_frames) { y.x *= Wf; y.y *= Wf; }
      if (a.PFW != nullptr && valid) a.PFW[(size_t)tl * a.Gp + g] = Wf;
    } else if constexpr (PF == 1) {
      // ZelinskiFilter_f (postfilter.cc:57-140); alpha = 0 for the first two frames (postfilter.cc:460-463)
      const float al = (t >= 2) ? a.pf_alpha : 0.f;
      float2 s = make_float2(0.f, 0.f);
      float den = 0.f;
      if constexpr (PK) zelinski_csd_step<C, true>(x, ta, csd, psd, al, live, s, den);   // btkb_nlms_math.cuh (its PK = false branch is this code)
      else {
      float2 z[C];
#pragma unroll
      for (int c = 0; c < C; c++) z[c] = cmulc(x[c], ta[c]);
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
      }
      float num = (a.pf_type & 1) ? fmaxf(s.x, 0.f) : sqrtf(fmaf(s.x, s.x, s.y * s.y));
      float Wf = (num / den) * (2.0f / ((float)C - 1.0f));
      Wf = (Wf >= 1.0f) ? 1.0f : Wf;
      Wf = (Wf < 1.0e-4f) ? 1.0e-4f : Wf;
      if (!(den > 0.f)) Wf = 1.0e-4f;  // all-zero snapshot: the reference computes 0/0 = NaN -> fails both tests; keep finite
      if (t - 1 >= a.pf_min_frames) { y.x *= Wf; y.y *= Wf; }
      if (a.PFW != nullptr && valid) a.PFW[(size_t)tl * a.Gp + g] = Wf;
    }

    if (valid) a.Y[(size_t)tl * a.Gp + g] = live ? y : make_float2(0.f, 0.f);
  }

  if (MODE == MODE_LMS && valid) {
    if (a.UA != nullptr) {
#pragma unroll
      for (int c = 0; c < C; c++) a.UA[(size_t)c * a.Gp + g] = uw[c];
    }
    if (k == 0 && a.stats_updates != nullptr) a.stats_updates[u] = (float)n_updates;
  }
  if (has_st) {
    const size_t gp = (size_t)a.Gp;
    if (MODE == MODE_LMS) { stp[0 * gp] = se; stp[1 * gp] = Eavg; stp[2 * gp] = gamma; stp[3 * gp] = __int_as_float(slow_cnt); stp[4 * gp] = __int_as_float(n_updates); }
    if (PF == 1) {
#pragma unroll
      for (int i = 0; i < NP; i++) { stp[(8 + 2 * i) * gp] = csd[i].x; stp[(9 + 2 * i) * gp] = csd[i].y; }
    }
    if (PF >= 2) { stp[8 * gp] = S1.x; stp[9 * gp] = S1.y; stp[10 * gp] = S2.x; stp[11 * gp] = S2.y; }
    if (PF) {
#pragma unroll
      for (int c = 0; c < C; c++) stp[((PF == 1 ? 8 + 2 * NP : 12) + c) * gp] = psd[c];
    }
  }
}

// K2: R[u][k] += x x^H over the frames flagged in noise_mask (pybeamformer.py:976-982)
// ---------------------------------------------------------------------------------------------------------------------
// RLS sidelobe canceller: SubbandGSCRLSBeamformer.__iter__ (lib/pybeamformer.py:817-901), one thread per (utterance, bin)
// chain, in the C-dimensional projector form (oracle/restate.py gsc_rls_projector; equal to the reference's B-form to
// 1e-15 in fp64).  With Q = conj(B) (Q^H Q = I, Q^H v = 0) the kernel carries u = waH Q^H (C complex) and the Hermitian
// Pt = Q Pz Q^H (C real diagonals + C(C-1)/2 complex) in registers, so the (C-1) x C blocking product never runs:
//   x~ = x - C Yc v        p = Pt x~        ip = Re(x~^H p)        Pt <- (Pt - p p^H / (mu + ip)) / mu
//   u  <- u + gamma ep p^H / (mu + ip) - reg u Pt^H      (ep = Yc - u.x with the OLD u, pybeamformer.py:843-847)
//   |u|^2 > alpha2: quadratic constraint with va = Pt u^H (:851-861);  |u|^2 > max_wa_l2norm: rescale u and reset
//   Pt = (I - C v v^H) / init_diagonal_load (:862-865).
// PK = true (the default; BTKB_PERBIN_PACKED=0 for the scalar form): the always-executed part of the update (projector step, p = Pt x~, the rank-one
// update of Pt, the new u with its regularisation term) in packed 2 x fp32 instructions (rls_core_step, btkb_nlms_math.cuh; bit-identical,
// checked on the CPU); the rarely taken constraint branches stay scalar.
template <int C, bool PK = false>
__global__ void __launch_bounds__(TILE) k_perbin_rls(const __grid_constant__ CUtensorMap tmX, PerBinArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  const int k = valid ? g - u * a.K : 0;
  const int len = a.lengths ? a.lengths[u] : 0;
  const int Tu = valid ? (a.tu ? a.tu[u] - a.t_base : frames_of(len, a.D, a.laN, a.pdA)) : 0;
  const bool has_st = a.ST != nullptr && valid;   // carried state of a streamed chunk (see k_perbin)
  float* const stp = a.ST + (valid ? g : 0);
  const bool st_load = has_st && a.st_load != 0;

  TileRing<C> ring;
  ring.init(smem_raw, &tmX, g0, a.T);

  float2 w[C], uw[C];
#pragma unroll
  for (int c = 0; c < C; c++) { w[c] = __ldg(a.W + (size_t)c * a.Gp + g); uw[c] = make_float2(0.f, 0.f); }
  const float inv_load = 1.0f / a.rls.init_diagonal_load;
  HermP<C> P;
  auto reset_P = [&]() {
#pragma unroll
    for (int i = 0; i < C; i++) {
      P.d[i] = (1.0f - (float)C * fmaf(w[i].x, w[i].x, w[i].y * w[i].y)) * inv_load;
#pragma unroll
      for (int j = 0; j < i; j++) {  // -(C v_i conj(v_j)) / load
        P.o[HermP<C>::idx(i, j)] = make_float2(-(float)C * fmaf(w[i].x, w[j].x, w[i].y * w[j].y) * inv_load,
                                               -(float)C * fmaf(w[i].y, w[j].x, -w[i].x * w[j].y) * inv_load);
      }
    }
  };
  reset_P();
  float Eavg = a.rls.init_diagonal_load;
  int n_updates = 0;
  float e_next = (a.T > 0) ? __ldg(a.E + u) : 0.f;
  const float one_m_beta = 1.0f - a.rls.beta;
  const float inv_sil = 1.0f / a.rls.sil_thresh;
  const float mu = a.rls.mu, inv_mu = 1.0f / a.rls.mu;
  const int opt = a.rls.constraint_option;
  if (st_load) {
    const size_t gp = (size_t)a.Gp;
    Eavg = stp[0 * gp]; n_updates = __float_as_int(stp[1 * gp]);
#pragma unroll
    for (int c = 0; c < C; c++) { uw[c] = a.UA[(size_t)c * a.Gp + g]; P.d[c] = stp[(8 + c) * gp]; }
#pragma unroll
    for (int i = 0; i < C * (C - 1) / 2; i++) P.o[i] = make_float2(stp[(8 + C + 2 * i) * gp], stp[(9 + C + 2 * i) * gp]);
  }

  for (int tl = 0; tl < a.T; tl++) {
    const int t = a.t_base + tl;   // absolute frame number
    float2 x[C];
    ring.fetch(tl, a.T, x);
    const float energy = e_next;
    if (tl + 1 < a.T) e_next = __ldg(a.E + (size_t)(tl + 1) * a.U + u);
    float2 y = cdot<C, true, PK>(x, w);   // Yc = v^H x
    const bool live = tl < Tu;
    const bool adapt = energy > (Eavg * inv_sil);
    if (adapt && live) {
      float2 un[C];
      rls_core_step<C, PK>(x, w, uw, y, P, mu, inv_mu, a.rls.gamma, a.rls.regularization_param, un);
      if (opt > 0) {
        float n2 = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) n2 = fmaf(un[c].x, un[c].x, fmaf(un[c].y, un[c].y, n2));
        if ((opt == 1 || opt == 3) && n2 > a.rls.alpha2) {
          float2 va[C];
          float qa = 0.f, qb = 0.f;
#pragma unroll
          for (int i = 0; i < C; i++) {  // va = Pt conj(un)
            float2 s = make_float2(P.d[i] * un[i].x, -P.d[i] * un[i].y);
#pragma unroll
            for (int j = 0; j < C; j++) {
              if (j == i) continue;
              const float2 pij = P.get(i, j);
              cmac_conj(s, pij, un[j]);
            }
            va[i] = s;
            qa = fmaf(s.x, s.x, fmaf(s.y, s.y, qa));
            qb = fmaf(s.x, un[i].x, fmaf(-s.y, un[i].y, qb));   // Re(conj(va_i) conj(un_i))
          }
          const float b = -2.0f * qb, cc = n2 - a.rls.alpha2;
          const float arg = fmaf(b, b, -4.0f * qa * cc);
          const float betaK = (arg > 0.f) ? -(b + sqrtf(arg)) / (2.0f * qa) : -b / (2.0f * qa);
#pragma unroll
          for (int c = 0; c < C; c++) { un[c].x = fmaf(-betaK, va[c].x, un[c].x); un[c].y = fmaf(betaK, va[c].y, un[c].y); }
        }
        if (opt >= 2 && n2 > a.rls.max_wa_l2norm) {
          const float sc = sqrtf(a.rls.max_wa_l2norm / n2);
#pragma unroll
          for (int c = 0; c < C; c++) { un[c].x *= sc; un[c].y *= sc; }
          reset_P();
        }
      }
#pragma unroll
      for (int c = 0; c < C; c++) uw[c] = un[c];
      if (k == 0) n_updates++;
    }
    if (t >= a.rls.min_frames) y = csub(y, cdot<C, false, PK>(uw, x));
    Eavg = fmaf(Eavg, a.rls.beta, one_m_beta * energy);
    if (valid) a.Y[(size_t)tl * a.Gp + g] = live ? y : make_float2(0.f, 0.f);
  }
  if (valid) {
    if (a.UA != nullptr) {
#pragma unroll
      for (int c = 0; c < C; c++) a.UA[(size_t)c * a.Gp + g] = uw[c];
    }
    if (k == 0 && a.stats_updates != nullptr) a.stats_updates[u] = (float)n_updates;
  }
  if (has_st) {
    const size_t gp = (size_t)a.Gp;
    stp[0 * gp] = Eavg; stp[1 * gp] = __int_as_float(n_updates);
#pragma unroll
    for (int c = 0; c < C; c++) stp[(8 + c) * gp] = P.d[c];
#pragma unroll
    for (int i = 0; i < C * (C - 1) / 2; i++) { stp[(8 + C + 2 * i) * gp] = P.o[i].x; stp[(9 + C + 2 * i) * gp] = P.o[i].y; }
  }
}

template <int C, bool PK = false>   // PK: x_i conj(x_j) and its accumulation as 3 packed instructions instead of 6 (BTKB_PERBIN_PACKED=1)
__global__ void __launch_bounds__(TILE) k_covariance(const __grid_constant__ CUtensorMap tmX, PerBinArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  TileRing<C> ring;
  ring.init(smem_raw, &tmX, g0, a.T);
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
    ring.fetch(t, a.T, x);
    const unsigned char mk = m_next;
    if (t + 1 < a.T) m_next = a.noise_mask[(size_t)(t + 1) * a.U + u];
    if (mk) {
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C; i++) {
        dg[i] = fmaf(x[i].x, x[i].x, fmaf(x[i].y, x[i].y, dg[i]));
#pragma unroll
        for (int j = i + 1; j < C; j++) {
          if constexpr (PK) off[idx] = f2_add(off[idx], f2_cmulc(x[i], x[j]));
          else { float2 p = cmulc(x[i], x[j]); off[idx].x += p.x; off[idx].y += p.y; }
          idx++;
        }
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

// SpectralMatrixArray::update (beamformer.cc:122-143) over all resident frames: R <- mu R + (1 - mu) x x^T (NOCONJ, the
// reference's arithmetic: complex-symmetric) or x x^H (Hermitian).  Upper triangle incl. diagonal in registers.
template <int C, int NOCONJ>
__global__ void __launch_bounds__(TILE) k_spectral_recursion(const __grid_constant__ CUtensorMap tmX, PerBinArgs a, float mu) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  const int Tu = valid ? frames_of(a.lengths[u], a.D, a.laN, a.pdA) : 0;
  TileRing<C> ring;
  ring.init(smem_raw, &tmX, g0, a.T);
  constexpr int NU = C * (C + 1) / 2;
  float2 r[NU];
#pragma unroll
  for (int i = 0; i < NU; i++) r[i] = make_float2(0.f, 0.f);
  const float om = 1.0f - mu;
  for (int t = 0; t < a.T; t++) {
    float2 x[C];
    ring.fetch(t, a.T, x);
    if (t < Tu) {
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C; i++)
#pragma unroll
        for (int j = i; j < C; j++) {
          const float2 p = NOCONJ ? cmul(x[i], x[j]) : cmulc(x[i], x[j]);
          r[idx].x = fmaf(mu, r[idx].x, om * p.x); r[idx].y = fmaf(mu, r[idx].y, om * p.y);
          idx++;
        }
    }
  }
  if (valid) {
    int idx = 0;
#pragma unroll
    for (int i = 0; i < C; i++)
#pragma unroll
      for (int j = i; j < C; j++) {
        a.R[(size_t)(i * C + j) * a.Gp + g] = r[idx];
        if (j != i) a.R[(size_t)(j * C + i) * a.Gp + g] = NOCONJ ? r[idx] : make_float2(r[idx].x, -r[idx].y);
        idx++;
      }
  }
}

template <int C>
static size_t ring_smem() { return ring_bytes<C>(); }


template <int C>
static cudaError_t launch_perbin_c(const PerBinArgs& a, cudaStream_t st) {
  const size_t smem = ring_smem<C>();
  const int grid = (a.G + TILE - 1) / TILE;
  CUtensorMap tm;
  { cudaError_t e = make_tensor_map(&tm, a, C); if (e != cudaSuccess) return e; }
  const bool lms = a.kind == BTKB_BF_GSC_LMS;
  const int pf = a.pf_kind;
  constexpr int NPAIR = C * (C - 1) / 2;
#define BTKB_LAUNCH(MODE_, PF_)                                                                  \
  do {                                                                                           \
    auto kern = k_perbin<C, MODE_, PF_>;                                                         \
    const size_t sm = smem + ((PF_) == 3 ? 2 * (NPAIR + C) : (PF_) == 2 ? (NPAIR + C) : 0) * TILE * sizeof(float2); \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    if (e != cudaSuccess) return e;                                                              \
    kern<<<grid, TILE, sm, st>>>(tm, a);                                                         \
    return cudaGetLastError();                                                                   \
  } while (0)
  if (a.kind == BTKB_BF_GSC_RLS) {
    if (pf != BTKB_PF_NONE) return cudaErrorInvalidValue;
    auto kern = env_packed("BTKB_PERBIN_PACKED") ? k_perbin_rls<C, true> : k_perbin_rls<C, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TILE, smem, st>>>(tm, a);
    return cudaGetLastError();
  }
  if (lms) {
    if (pf != BTKB_PF_NONE) return cudaErrorInvalidValue;
    if (env_packed("BTKB_PERBIN_PACKED")) {   // packed 2 x fp32 NLMS recurrence (bit-identical) unless BTKB_PERBIN_PACKED=0; read at every launch
      auto kern = k_perbin<C, MODE_LMS, 0, true>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kern<<<grid, TILE, smem, st>>>(tm, a);
      return cudaGetLastError();
    }
    BTKB_LAUNCH(MODE_LMS, 0);
  }
  if (pf == BTKB_PF_ZELINSKI) {
    if (env_packed("BTKB_PERBIN_PACKED")) {   // packed 2 x fp32 CSD recursions (bit-identical)
      auto kern = k_perbin<C, MODE_STATIC, 1, true>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kern<<<grid, TILE, smem, st>>>(tm, a);
      return cudaGetLastError();
    }
    BTKB_LAUNCH(MODE_STATIC, 1);
  }
  const bool pk = env_packed("BTKB_PERBIN_PACKED");
#define BTKB_LAUNCH_PK(PF_)                                                                      \
  do {                                                                                           \
    auto kern = k_perbin<C, MODE_STATIC, PF_, true>;                                             \
    const size_t sm = smem + ((PF_) == 3 ? 2 * (NPAIR + C) : (NPAIR + C)) * TILE * sizeof(float2); \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    if (e != cudaSuccess) return e;                                                              \
    kern<<<grid, TILE, sm, st>>>(tm, a);                                                         \
    return cudaGetLastError();                                                                   \
  } while (0)
  if (pf == BTKB_PF_MCCOWAN) { if (!a.PFQ) return cudaErrorInvalidValue; if (pk) BTKB_LAUNCH_PK(2); BTKB_LAUNCH(MODE_STATIC, 2); }
  if (pf == BTKB_PF_LEFKIMMIATIS) { if (!a.PFQ || !a.LAM) return cudaErrorInvalidValue; if (pk) BTKB_LAUNCH_PK(3); BTKB_LAUNCH(MODE_STATIC, 3); }
#undef BTKB_LAUNCH_PK
  BTKB_LAUNCH(MODE_STATIC, 0);
#undef BTKB_LAUNCH
}

#if BTKB_CSET == 0
cudaError_t launch_perbin(const PerBinArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  switch (a.C) {
    case 2: return launch_perbin_c<2>(a, st);
    case 4: return launch_perbin_c<4>(a, st);
    case 8: return launch_perbin_c<8>(a, st);
    default: return launch_perbin_cset1(a, st);
  }
}
#else
cudaError_t launch_perbin_cset1(const PerBinArgs& a, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_perbin_c<1>(a, st);
    case 3: return launch_perbin_c<3>(a, st);
    case 5: return launch_perbin_c<5>(a, st);
    case 6: return launch_perbin_c<6>(a, st);
    case 7: return launch_perbin_c<7>(a, st);
    default: return cudaErrorInvalidValue;
  }
}
#endif

template <int C>
static cudaError_t launch_cov_c(const PerBinArgs& a, cudaStream_t st) {
  const size_t smem = ring_smem<C>();
  auto kern = env_packed("BTKB_PERBIN_PACKED") ? k_covariance<C, true> : k_covariance<C, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  CUtensorMap tm;
  e = make_tensor_map(&tm, a, C);
  if (e != cudaSuccess) return e;
  kern<<<(a.G + TILE - 1) / TILE, TILE, smem, st>>>(tm, a);
  return cudaGetLastError();
}

#if BTKB_CSET == 0
cudaError_t launch_covariance(const PerBinArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  switch (a.C) {
    case 2: return launch_cov_c<2>(a, st);
    case 4: return launch_cov_c<4>(a, st);
    case 8: return launch_cov_c<8>(a, st);
    default: return launch_covariance_cset1(a, st);
  }
}
#else
cudaError_t launch_covariance_cset1(const PerBinArgs& a, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_cov_c<1>(a, st);
    case 3: return launch_cov_c<3>(a, st);
    case 5: return launch_cov_c<5>(a, st);
    case 6: return launch_cov_c<6>(a, st);
    case 7: return launch_cov_c<7>(a, st);
    default: return cudaErrorInvalidValue;
  }
}
#endif

}  // namespace btkb

namespace btkb {
template <int C>
static cudaError_t launch_rec_c(const PerBinArgs& a, float mu, int noconj, cudaStream_t st) {
  const size_t smem = ring_smem<C>();
  CUtensorMap tm;
  cudaError_t e = make_tensor_map(&tm, a, C);
  if (e != cudaSuccess) return e;
  const int grid = (a.G + TILE - 1) / TILE;
  if (noconj) {
    auto kern = k_spectral_recursion<C, 1>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TILE, smem, st>>>(tm, a, mu);
  } else {
    auto kern = k_spectral_recursion<C, 0>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TILE, smem, st>>>(tm, a, mu);
  }
  return cudaGetLastError();
}
#if BTKB_CSET == 0
cudaError_t launch_spectral_recursion(const PerBinArgs& a, float mu, int noconj, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  switch (a.C) {
    case 2: return launch_rec_c<2>(a, mu, noconj, st);
    case 4: return launch_rec_c<4>(a, mu, noconj, st);
    case 8: return launch_rec_c<8>(a, mu, noconj, st);
    default: return launch_spectral_recursion_cset1(a, mu, noconj, st);
  }
}
#else
cudaError_t launch_spectral_recursion_cset1(const PerBinArgs& a, float mu, int noconj, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_rec_c<1>(a, mu, noconj, st);
    case 3: return launch_rec_c<3>(a, mu, noconj, st);
    case 5: return launch_rec_c<5>(a, mu, noconj, st);
    case 6: return launch_rec_c<6>(a, mu, noconj, st);
    case 7: return launch_rec_c<7>(a, mu, noconj, st);
    default: return cudaErrorInvalidValue;
  }
}
#endif
}  // namespace btkb
