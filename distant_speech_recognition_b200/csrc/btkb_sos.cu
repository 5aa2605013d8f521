// btkb_sos.cu — second-order-statistics batch beamformers: blind MVDR (MMSE) and GEV (sm_100a).
//
// Replaces the per-frame / per-bin Python loops of lib/pybeamformer.py (reference: btk20_src/):
//   SubbandSOSBatchBeamformer.accu_stats_from_label    pybeamformer.py:1063-1127   target / noise covariance from a VAD label
//   SubbandSOSBatchBeamformer.accu_stats_from_tfmask   pybeamformer.py:1129-1183   ... from time-frequency masks
//   SubbandBlindMVDRBeamformer.finalize_stats / calc_beamformer_weights  :1257-1295   w = Rn^-1 Rt u / (offset + tr(Rn^-1 Rt))
//   SubbandGEVBeamformer.finalize_stats / calc_beamformer_weights        :1311-1357   principal generalised eigenvector of
//                                                                                     (Rt, Rn), phase-aligned bin to bin
//   improve_matrix_condition                           pybeamformer.py:1231-1240
// The weight-apply (SubbandSOSBatchBeamformer.__iter__, :1191-1207: y = wqH . x for every bin incl. DC) is the static mode of
// the per-bin kernel (btkb_perbin.cu) with W = conj(wqH).
//
//   k_sos_frame_weights  per utterance: the reference's label walk (running elapsed time, segment cursor) x energy gate
//   k_sos_scatter_mask   host-layout masks [U][Tm][K] -> [T][Gp] (chain-fastest, coalesced for the covariance kernel)
//   k_sos_cov<C>         one thread per (utterance, bin) chain, X through the tensor-map TMA ring; blockIdx.y = target / noise;
//                        Hermitian accumulators in fp64 registers (products of the fp32 snapshots are exact in fp64), counts
//                        with the reference's integer-array truncation (int_array[m] += float, :1157-1162)
//   k_sos_solve<C>       fp64, one thread per chain: normalisation, loading, LU solve (blind MVDR) or Cholesky reduction +
//                        cyclic complex Jacobi + back-substitution (GEV)
//   k_sos_align          per utterance: the bin-to-bin phase alignment (:1339-1341) is a serial scan over bins
#include "btkb_tile_ring.cuh"
#include "btkb_sos_math.cuh"
#include "../../include/btkb.h"
#include <math.h>

namespace btkb {
namespace {

__global__ void k_sos_frame_weights(SosArgs a) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= a.U) return;
  const int Tu = frames_of(a.lengths[u], a.D, a.laN, a.pdA);
  double elapsed = 0.0;
  const double dt = (double)a.D / (double)a.samplerate;
  int labx = 0;
  const double* lab = a.labels ? a.labels + (size_t)u * a.NL * 2 : nullptr;
  float* wt = a.wtu + (size_t)u;
  float* wn = a.wtu + (size_t)a.T * a.U + u;
  for (int t = 0; t < a.T; t++) {
    bool is_target = false;
    if (lab != nullptr && labx < a.NL) {   // pybeamformer.py:1086-1091
      const double s = lab[2 * labx], e = lab[2 * labx + 1];
      if (elapsed >= s && (elapsed <= e || e < 0)) is_target = true;
      else if (elapsed > e) labx += 1;
    }
    const bool gate = (t < Tu) && (a.E[(size_t)t * a.U + u] > a.thr);
    float ft, fn;
    if (lab != nullptr) { ft = (gate && is_target) ? 1.f : 0.f; fn = (gate && !is_target) ? 1.f : 0.f; }
    else { ft = fn = gate ? 1.f : 0.f; }   // TF-mask mode: the masks decide, the energy gate applies to both
    wt[(size_t)t * a.U] = ft; wn[(size_t)t * a.U] = fn;
    elapsed += dt;
  }
}

// src [U][Tm][K] (host layout, already on the device) -> dst [T][Gp]; frames beyond Tm read as 0
__global__ void k_sos_scatter_mask(const float* __restrict__ src, float* __restrict__ dst, int U, int Tm, int T, int K, int Gp) {
  const size_t total = (size_t)T * U * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K); size_t r = i / K;
    const int u = (int)(r % U); const int t = (int)(r / U);
    dst[(size_t)t * Gp + (size_t)u * K + k] = (t < Tm) ? src[((size_t)u * Tm + t) * K + k] : 0.f;
  }
}

template <int C>
__global__ void __launch_bounds__(TILE) k_sos_cov(const __grid_constant__ CUtensorMap tmX, SosArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int set = blockIdx.y;   // 0 target, 1 noise
  const int g0 = blockIdx.x * TILE;
  const int g = g0 + threadIdx.x;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  TileRing<C> ring;
  ring.init(smem_raw, &tmX, g0, a.T);
  constexpr int NP = C * (C - 1) / 2;
  double offr[NP > 0 ? NP : 1], offi[NP > 0 ? NP : 1], dg[C];
#pragma unroll
  for (int i = 0; i < NP; i++) { offr[i] = 0.0; offi[i] = 0.0; }
#pragma unroll
  for (int c = 0; c < C; c++) dg[c] = 0.0;
  double cnt = 0.0;
  const float* wtu = a.wtu + (size_t)set * a.T * a.U + u;
  const float* msk = (set == 0) ? a.mask_t : a.mask_j;
  if (msk != nullptr) msk += valid ? g : 0;
  // frame weight and mask value are fetched one frame ahead so the loop never waits on them
  float w_next = (a.T > 0) ? wtu[0] : 0.f;
  float m_next = (msk != nullptr && a.T > 0) ? msk[0] : 1.f;
  for (int t = 0; t < a.T; t++) {
    float2 x[C];
    ring.fetch(t, a.T, x);
    float w = w_next;
    const float mv = m_next;
    if (t + 1 < a.T) { w_next = wtu[(size_t)(t + 1) * a.U]; if (msk != nullptr) m_next = msk[(size_t)(t + 1) * a.Gp]; }
    if (msk != nullptr) w = (mv > 0.f) ? w * mv : 0.f;   // `if mask[frame_no][m] > 0` (:1151-1162)
    if (w != 0.f && valid) {
      const double wd = (double)w;
      cnt = trunc(cnt + wd);    // numpy int array element += float: every addition truncates
      double xr[C], xi[C];
#pragma unroll
      for (int c = 0; c < C; c++) { xr[c] = (double)x[c].x * wd; xi[c] = (double)x[c].y * wd; }   // w x_i, then (w x_i) conj(x_j)
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C; i++) {
        dg[i] = fma(xr[i], (double)x[i].x, fma(xi[i], (double)x[i].y, dg[i]));
#pragma unroll
        for (int j = i + 1; j < C; j++) {
          const double br = (double)x[j].x, bi = (double)x[j].y;
          offr[idx] = fma(xr[i], br, fma(xi[i], bi, offr[idx]));
          offi[idx] = fma(xi[i], br, fma(-xr[i], bi, offi[idx]));
          idx++;
        }
      }
    }
  }
  if (valid) {
    double2* R = a.Rs + (size_t)set * C * C * a.Gp;
    double* cn = a.cnt + (size_t)set * a.Gp;
    const bool acc = a.accumulate != 0;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < C; i++) {
      double2* d = R + (size_t)(i * C + i) * a.Gp + g;
      *d = make_double2(dg[i] + (acc ? d->x : 0.0), 0.0);
#pragma unroll
      for (int j = i + 1; j < C; j++) {
        double2* q = R + (size_t)(i * C + j) * a.Gp + g;
        const double vr = offr[idx] + (acc ? q->x : 0.0), vi = offi[idx] + (acc ? q->y : 0.0);
        *q = make_double2(vr, vi);
        R[(size_t)(j * C + i) * a.Gp + g] = make_double2(vr, -vi);
        idx++;
      }
    }
    cn[g] = cnt + (acc ? cn[g] : 0.0);
  }
}

// error bits written to a.err: 1 no target statistics, 2 no noise statistics, 4 factorisation failed
template <int C>
__global__ void __launch_bounds__(64) k_sos_solve(SosArgs a, int kind, double gamma, int ref_micx, double offset) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.G) return;
  const double ct = a.cnt[g], cn = a.cnt[(size_t)a.Gp + g];
  if (!(ct > 0.0)) { atomicOr(a.err, 1); }
  if (!(cn > 0.0)) { atomicOr(a.err, 2); }
  if (!(ct > 0.0) || !(cn > 0.0)) {
    for (int c = 0; c < C; c++) { a.Wd[(size_t)c * a.Gp + g] = make_double2(0, 0); a.W[(size_t)c * a.Gp + g] = make_float2(0.f, 0.f); }
    return;
  }
  zd Rt[C][C], Rn[C][C];
  const double2* pt = a.Rs;
  const double2* pn = a.Rs + (size_t)C * C * a.Gp;
  const double st = (kind == BTKB_SOS_BMVDR) ? 1.0 / ct : 1.0;   // GEV skips the target normalisation (:1320-1322)
  const double sn = 1.0 / cn;
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      const double2 t = pt[(size_t)(i * C + j) * a.Gp + g], n = pn[(size_t)(i * C + j) * a.Gp + g];
      Rt[i][j] = zmk(t.x * st, t.y * st);
      Rn[i][j] = zmk(n.x * sn, n.y * sn);
    }
  zd w[C];
  const bool bad = !sos_solve_chain<C>(Rt, Rn, kind, gamma, ref_micx, offset, w);
  if (bad) {
    atomicOr(a.err, 4);
    for (int c = 0; c < C; c++) w[c] = zmk(0, 0);
  }
  for (int c = 0; c < C; c++) {
    a.Wd[(size_t)c * a.Gp + g] = make_double2(w[c].x, w[c].y);
    a.W[(size_t)c * a.Gp + g] = make_float2((float)w[c].x, (float)w[c].y);
  }
}

// wqH[m] *= exp(-j angle(inner(wqH[m], conj(wqH[m-1])))) for m = 1..K-1, each bin against the ALIGNED previous bin (:1339-1341)
__global__ void k_sos_align(SosArgs a) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= a.U) return;
  zd rot = zmk(1.0, 0.0);   // cumulative rotation of the previous bin
  for (int k = 1; k < a.K; k++) {
    const size_t g = (size_t)u * a.K + k;
    zd ip = zmk(0, 0);
    for (int c = 0; c < a.C; c++) {
      const double2 cur = a.Wd[(size_t)c * a.Gp + g], prv = a.Wd[(size_t)c * a.Gp + g - 1];
      ip = zadd(ip, zmulc(zmk(cur.x, cur.y), zmul(zmk(prv.x, prv.y), rot)));
    }
    const double m = sqrt(zabs2(ip));
    rot = (m > 0.0) ? zscale(zconj(ip), 1.0 / m) : zmk(1.0, 0.0);   // numpy.angle(0) = 0
    for (int c = 0; c < a.C; c++) {
      const double2 cur = a.Wd[(size_t)c * a.Gp + g];
      const zd v = zmul(zmk(cur.x, cur.y), rot);
      a.W[(size_t)c * a.Gp + g] = make_float2((float)v.x, (float)v.y);
    }
  }
}

template <int C>
cudaError_t launch_cov_c(const SosArgs& a, cudaStream_t st) {
  const size_t smem = ring_bytes<C>();
  auto kern = k_sos_cov<C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  PerBinArgs pa;
  pa.X = a.X; pa.Gp = a.Gp; pa.T = a.T;   // the fields make_tensor_map reads
  CUtensorMap tm;
  e = make_tensor_map(&tm, pa, C);
  if (e != cudaSuccess) return e;
  kern<<<dim3((a.G + TILE - 1) / TILE, 2), TILE, smem, st>>>(tm, a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_sos_scatter_mask(const float* src, float* dst, int U, int Tm, int T, int K, int Gp, cudaStream_t st) {
  k_sos_scatter_mask<<<148 * 8, 256, 0, st>>>(src, dst, U, Tm, T, K, Gp);
  return cudaGetLastError();
}

cudaError_t launch_sos_accumulate(const SosArgs& a, cudaStream_t st, int* launches) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  k_sos_frame_weights<<<(a.U + 63) / 64, 64, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  switch (a.C) {
    case 2: e = launch_cov_c<2>(a, st); break;
    case 3: e = launch_cov_c<3>(a, st); break;
    case 4: e = launch_cov_c<4>(a, st); break;
    case 5: e = launch_cov_c<5>(a, st); break;
    case 6: e = launch_cov_c<6>(a, st); break;
    case 7: e = launch_cov_c<7>(a, st); break;
    case 8: e = launch_cov_c<8>(a, st); break;
    default: return cudaErrorInvalidValue;
  }
  if (launches) *launches += 2;
  return e;
}

cudaError_t launch_sos_solve(const SosArgs& a, int kind, double gamma, int ref_micx, double offset, cudaStream_t st, int* launches) {
  if (a.G <= 0) return cudaSuccess;
  const int bs = 64, gs = (a.G + bs - 1) / bs;
  switch (a.C) {
    case 2: k_sos_solve<2><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    case 3: k_sos_solve<3><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    case 4: k_sos_solve<4><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    case 5: k_sos_solve<5><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    case 6: k_sos_solve<6><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    case 7: k_sos_solve<7><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    case 8: k_sos_solve<8><<<gs, bs, 0, st>>>(a, kind, gamma, ref_micx, offset); break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (launches) (*launches)++;
  if (kind == BTKB_SOS_GEV) {
    k_sos_align<<<(a.U + 63) / 64, 64, 0, st>>>(a);
    e = cudaGetLastError();
    if (launches) (*launches)++;
  }
  return e;
}

}  // namespace btkb
