// btkb_nlms_math.cuh — the complex arithmetic of the per-bin NLMS recurrence (k_perbin<C, LMS>, btkb_perbin.cu), in a header of its
// own so that tests/host/fft_packed_host.cc can run the scalar (PK = false) and the packed 2 x fp32 (PK = true, btkb_f2.cuh) forms on
// the CPU and require bit-identical results.  Reference: SubbandGSCLMSBeamformer.__iter__, btk20_src/lib/pybeamformer.py:659-734, in
// the O(C) projector form of SURVEY.md App. A.3.
#pragma once
#include <cuda_runtime.h>
#include "btkb_f2.cuh"
#include "btkb_fft.cuh"   // cmulc

namespace btkb {

// FMA-only complex multiply-accumulate (4 FFMA each, no separate adds)
__device__ __forceinline__ void cmac(float2& acc, float2 a, float2 b) {        // acc += a b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cmac_conj(float2& acc, float2 a, float2 b) {   // acc += a conj(b)
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.y, b.x, acc.y); acc.y = fmaf(-a.x, b.y, acc.y);
}
// sum_c a[c] b[c] (or a[c] conj(b[c])) with two independent accumulators (halves the dependent FMA chain)
// PK = true: each complex MAC is two FFMA2 (btkb_f2.cuh) instead of four FFMA; same operations and roundings per component.
template <int C, bool CONJ, bool PK = false>
__device__ __forceinline__ float2 cdot(const float2* a, const float2* b) {
  float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < C; c += 2) {
    if constexpr (PK) {
      if (CONJ) { s0 = f2_cmac_conj(s0, a[c], b[c]); if (c + 1 < C) s1 = f2_cmac_conj(s1, a[c + 1], b[c + 1]); }
      else { s0 = f2_cmac(s0, a[c], b[c]); if (c + 1 < C) s1 = f2_cmac(s1, a[c + 1], b[c + 1]); }
    } else {
    if (CONJ) { cmac_conj(s0, a[c], b[c]); if (c + 1 < C) cmac_conj(s1, a[c + 1], b[c + 1]); }
    else { cmac(s0, a[c], b[c]); if (c + 1 < C) cmac(s1, a[c + 1], b[c + 1]); }
    }
  }
  if constexpr (PK) return f2_add(s0, s1);
  return make_float2(s0.x + s1.x, s0.y + s1.y);
}

// One adaptation step (pybeamformer.py:689-716): e = Yc - u.x (a-priori error with the OLD u), a = gamma / sub-band energy,
// (Q x)_c = x_c - C Yc v_c,  u~_c = (1 - a reg) u_c + a e conj((Q x)_c);  returns the two partial sums of ||u~||^2.
template <int C, bool PK>
__device__ __forceinline__ void nlms_adapt_step(const float2* x, const float2* w, float2* uw, float2 y, float gamma, float sub, float reg, float& n20,
                                                float& n21) {
  const float2 ux = cdot<C, false, PK>(uw, x);
  if constexpr (PK) {
    const float2 epa = f2_sub(y, ux);
    const float alphaK = gamma / sub;
    const float2 cy = f2_scale(y, (float)C);
    const float2 ea = f2_scale(epa, alphaK);
    const float keep = (reg > 0.f) ? 1.0f - alphaK * reg : 1.0f;
    float2 nn = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C; c++) {
      const float2 q = f2_sub_cmul(x[c], cy, w[c]);
      const float2 un = f2_add_cmulc(f2_scale(uw[c], keep), ea, q);
      uw[c] = un;
      nn = f2_fma(un, un, nn);
    }
    n20 = nn.x; n21 = nn.y;
  } else {
    const float2 epa = make_float2(y.x - ux.x, y.y - ux.y);
    const float alphaK = gamma / sub;
    const float2 cy = make_float2((float)C * y.x, (float)C * y.y);
    const float2 ea = make_float2(epa.x * alphaK, epa.y * alphaK);
    const float keep = (reg > 0.f) ? 1.0f - alphaK * reg : 1.0f;
    n20 = 0.f; n21 = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) {
      // (Q x)_c = x_c - C Yc v_c ;  u~_c = (1 - a reg) u_c + a e conj(q_c)
      const float qx = fmaf(-cy.x, w[c].x, fmaf(cy.y, w[c].y, x[c].x));
      const float qy = fmaf(-cy.x, w[c].y, fmaf(-cy.y, w[c].x, x[c].y));
      const float unx = fmaf(ea.y, qy, fmaf(ea.x, qx, keep * uw[c].x));
      const float uny = fmaf(-ea.x, qy, fmaf(ea.y, qx, keep * uw[c].y));
      uw[c] = make_float2(unx, uny);
      n20 = fmaf(unx, unx, n20); n21 = fmaf(uny, uny, n21);
    }
  }
}

// One frame of the Zelinski post-filter's statistics (ZelinskiFilter_f, btk20_src/postfilter/postfilter.cc:57-140): time-aligned
// snapshot z_c = x_c conj(ta_c), recursive cross-spectral densities phi_ij <- al phi_ij + (1 - al) z_i conj(z_j) over the C(C-1)/2 pairs
// (summed into s), recursive power spectral densities (summed into den).  `live` is false past the end of the utterance: the state
// is then left untouched.
template <int C, bool PK>
__device__ __forceinline__ void zelinski_csd_step(const float2* x, const float2* ta, float2* csd, float* psd, float al, bool live, float2& s, float& den) {
  float2 z[C];
#pragma unroll
  for (int c = 0; c < C; c++) z[c] = PK ? f2_cmulc(x[c], ta[c]) : cmulc(x[c], ta[c]);
  s = make_float2(0.f, 0.f);
  den = 0.f;
  int idx = 0;
#pragma unroll
  for (int i = 0; i < C - 1; i++)
#pragma unroll
    for (int j = i + 1; j < C; j++) {
      if constexpr (PK) {
        const float2 zz = f2_cmulc(z[i], z[j]);
        const float2 ph = (al > 0.f) ? f2_fma_s(csd[idx], al, f2_scale(zz, 1.f - al)) : zz;
        if (live) csd[idx] = ph;
        s = f2_add(s, ph);
      } else {
      float2 zz = cmulc(z[i], z[j]);
      float2 ph = (al > 0.f) ? make_float2(fmaf(al, csd[idx].x, (1.f - al) * zz.x), fmaf(al, csd[idx].y, (1.f - al) * zz.y)) : zz;
      if (live) csd[idx] = ph;
      s.x += ph.x; s.y += ph.y;
      }
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

// ---- RLS sidelobe canceller (k_perbin_rls, btkb_perbin.cu; SubbandGSCRLSBeamformer.__iter__, lib/pybeamformer.py:817-901) ----
template <int C>
struct HermP {   // Hermitian C x C: d[i] real diagonal, o[idx(i,j)], i > j, lower triangle
  float d[C];
  float2 o[C * (C - 1) / 2 > 0 ? C * (C - 1) / 2 : 1];
  __device__ __forceinline__ static constexpr int idx(int i, int j) { return i * (i - 1) / 2 + j; }  // i > j
  __device__ __forceinline__ float2 get(int i, int j) const {
    if (i == j) return make_float2(d[i], 0.f);
    if (i > j) return o[idx(i, j)];
    const float2 t = o[idx(j, i)];
    return make_float2(t.x, -t.y);
  }
};

// The always-executed part of one RLS adaptation step in the projector form (pybeamformer.py:835-849):
//   x~ = x - C Yc v,  p = Pt x~,  ip = Re(x~^H p),  Pt <- (Pt - p p^H / (mu + ip)) / mu,
//   un = u + gamma ep conj(p) / (mu + ip) - reg u Pt^H   (ep = Yc - u.x with the OLD u, the NEW Pt in the regularisation term).
// The scalar form spells out the multiply-adds nvcc contracts (a - b c -> fma(b, -c, a)), so that the packed form — whose operations
// are explicit instructions — rounds the same way on the device; the real chain `ip` and the real diagonal of Pt stay scalar in both.
template <int C, bool PK>
__device__ __forceinline__ void rls_core_step(const float2* x, const float2* w, const float2* uw, float2 y, HermP<C>& P, float mu, float inv_mu,
                                                       float gamma, float reg, float2* un) {
  float2 xt[C], pv[C];
  if constexpr (PK) {
    const float2 cy = f2_scale(y, (float)C);
#pragma unroll
    for (int c = 0; c < C; c++) xt[c] = f2_sub_cmul(x[c], cy, w[c]);
#pragma unroll
    for (int i = 0; i < C; i++) {
      float2 s = f2_scale(xt[i], P.d[i]);
#pragma unroll
      for (int j = 0; j < C; j++) {
        if (j == i) continue;
        if (j < i) s = f2_cmac(s, P.o[HermP<C>::idx(i, j)], xt[j]);
        else s = f2_cmac_conj(s, xt[j], P.o[HermP<C>::idx(j, i)]);   // conj(P_ji) x_j
      }
      pv[i] = s;
    }
  } else {
    const float2 cy = make_float2((float)C * y.x, (float)C * y.y);
#pragma unroll
    for (int c = 0; c < C; c++)
      xt[c] = make_float2(fmaf(-cy.x, w[c].x, fmaf(cy.y, w[c].y, x[c].x)), fmaf(-cy.x, w[c].y, fmaf(-cy.y, w[c].x, x[c].y)));
#pragma unroll
    for (int i = 0; i < C; i++) {
      float2 s = make_float2(P.d[i] * xt[i].x, P.d[i] * xt[i].y);
#pragma unroll
      for (int j = 0; j < C; j++) {
        if (j == i) continue;
        if (j < i) cmac(s, P.o[HermP<C>::idx(i, j)], xt[j]);
        else cmac_conj(s, xt[j], P.o[HermP<C>::idx(j, i)]);   // conj(P_ji) x_j
      }
      pv[i] = s;
    }
  }
  float ip = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) ip = fmaf(xt[c].x, pv[c].x, fmaf(xt[c].y, pv[c].y, ip));
  const float inv = 1.0f / (mu + ip);
  // Pt <- (Pt - p p^H inv) / mu
#pragma unroll
  for (int i = 0; i < C; i++) {
    P.d[i] = fmaf(fmaf(pv[i].x, pv[i].x, pv[i].y * pv[i].y), -inv, P.d[i]) * inv_mu;
#pragma unroll
    for (int j = 0; j < i; j++) {
      float2& e = P.o[HermP<C>::idx(i, j)];
      if constexpr (PK) {
        e = f2_scale(f2_fma_s(f2_cmulc(pv[i], pv[j]), -inv, e), inv_mu);
      } else {
        const float2 pp = cmulc(pv[i], pv[j]);  // p_i conj(p_j)
        e.x = fmaf(pp.x, -inv, e.x) * inv_mu; e.y = fmaf(pp.y, -inv, e.y) * inv_mu;
      }
    }
  }
  const float2 ux = cdot<C, false, PK>(uw, x);
  const float gi = gamma * inv;
  if constexpr (PK) {
    const float2 ge = f2_scale(f2_sub(y, ux), gi);
#pragma unroll
    for (int c = 0; c < C; c++) un[c] = f2_cmac_conj(uw[c], ge, pv[c]);   // u + gamma ep conj(p) inv
  } else {
    const float2 ep = make_float2(y.x - ux.x, y.y - ux.y);
    const float2 ge = make_float2(gi * ep.x, gi * ep.y);
#pragma unroll
    for (int c = 0; c < C; c++) {
      un[c] = uw[c];
      cmac_conj(un[c], ge, pv[c]);
    }
  }
  if (reg > 0.f) {
#pragma unroll
    for (int c = 0; c < C; c++) {  // (u Pt^H)_c = sum_j u_j conj(Pt_cj)
      if constexpr (PK) {
        float2 s = f2_scale(uw[c], P.d[c]);
#pragma unroll
        for (int j = 0; j < C; j++) {
          if (j == c) continue;
          if (j < c) s = f2_cmac_conj(s, uw[j], P.o[HermP<C>::idx(c, j)]);
          else s = f2_cmac(s, uw[j], P.o[HermP<C>::idx(j, c)]);         // conj(P_cj) = P_jc
        }
        un[c] = f2_fma_s(s, -reg, un[c]);
      } else {
        float2 s = make_float2(uw[c].x * P.d[c], uw[c].y * P.d[c]);
#pragma unroll
        for (int j = 0; j < C; j++) {
          if (j == c) continue;
          if (j < c) cmac_conj(s, uw[j], P.o[HermP<C>::idx(c, j)]);
          else cmac(s, uw[j], P.o[HermP<C>::idx(j, c)]);         // conj(P_cj) = P_jc
        }
        un[c].x = fmaf(-reg, s.x, un[c].x); un[c].y = fmaf(-reg, s.y, un[c].y);
      }
    }
  }
}

}  // namespace btkb
