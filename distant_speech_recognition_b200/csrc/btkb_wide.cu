// btkb_wide.cu — per-bin kernels for wide arrays (C = 16, 32, 64 channels; configs[3] uses 64), sm_100a.
//
// Same mathematics and reference citations as btkb_perbin.cu (SubbandDS/GSC/MVDR apply, beamformer.cc:1095-1316,2537-2773;
// NLMS, lib/pybeamformer.py:659-734; SMI covariance, pybeamformer.py:948-1000; MVDR solve, beamformer.cc:2350-2402), but a
// (utterance, bin) chain is spread over L = C/8 adjacent lanes, 8 channels per lane, so all per-chain state still lives
// in registers; the per-bin complex reductions (Yc = v^H x, u.x, ||x||^2, ||u||^2) are finished with warp shuffles
// across the L lanes.  A CTA owns 16 consecutive chains: its mic x bin tile per frame is a [C rows][16 chains] box
// (128 B rows) fetched by one tensor-map TMA instruction with the 128-byte swizzle, which spreads the L lanes of a chain
// (reading the same column of 8 different rows) over different shared-memory banks.
//
// The covariance / MVDR-solve kernels here are the plain CUDA-core versions (one CTA per chain; C = 16, 32); the 64-mic
// covariance runs on the tensor cores instead (k_covariance_tc, btkb_cov_tc.cu: tcgen05 + TMEM).
#include <cuda.h>
#include "btkb_internal.h"
#include <algorithm>
#include "btkb_tensor_map.h"
#include "btkb_fft.cuh"
#include "../../include/btkb.h"
#include <cstdlib>

namespace btkb {
namespace wide {

constexpr int TC = 16;      // chains per CTA (16 x 8 B = one 128 B swizzle row)
constexpr int WSTAGES = 4;  // ring slots, one frame each

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* tm, int c0, int r0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(s32(dst)), "l"(tm), "r"(c0), "r"(r0), "r"(s32(bar)) : "memory");
}

// Butterfly sum over the L lanes of one chain.  `gm` names exactly those lanes: chains sharing a warp may diverge (padding
// chains, ragged utterance lengths, per-utterance silence gate), so a full-warp mask inside the adaptation branch would
// deadlock; the L lanes of one chain always take the same path.
template <int L>
__device__ __forceinline__ float red(float v, unsigned gm) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o);
  return v;
}
template <int L>
__device__ __forceinline__ float2 red2(float2 v, unsigned gm) { return make_float2(red<L>(v.x, gm), red<L>(v.y, gm)); }

__device__ __forceinline__ void mac(float2& acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x); acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void macc(float2& acc, float2 a, float2 b) {  // a conj(b)
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x); acc.y = fmaf(a.y, b.x, acc.y); acc.y = fmaf(-a.x, b.y, acc.y);
}

// MODE 0: static weights; MODE 1: NLMS
// PK = true (BTKB_PERBIN_PACKED=1, off by default): the complex MACs, the projector step and the update as FFMA2 pairs (btkb_f2.cuh);
// the two norm chains keep their scalar single-accumulator order, so the results stay bit-identical.
template <int L, int MODE, bool PK = false>
__global__ void __launch_bounds__(TC* L) k_perbin_wide(const __grid_constant__ CUtensorMap tmX, PerBinArgs a) {
  constexpr int C = 8 * L;
  constexpr int NTH = TC * L;
  constexpr int NWARPS = (NTH + 31) / 32;
  constexpr uint32_t SLOT = C * TC * sizeof(float2);   // C rows of 128 B
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stage = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)SLOT * WSTAGES);
  uint64_t* empty = full + WSTAGES;

  const int tid = threadIdx.x, j = tid / L, l = tid % L;
  const unsigned gm = ((L >= 32) ? 0xffffffffu : ((1u << L) - 1u)) << (((tid & 31) / L) * L);
  const int g0 = blockIdx.x * TC, g = g0 + j;
  const bool valid = g < a.G;
  const int u = valid ? g / a.K : a.U - 1;
  const int k = valid ? g - u * a.K : 0;
  const int Tu = valid ? frames_of(a.lengths[u], a.D, a.laN, a.pdA) : 0;
  const int T = a.T;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int s = 0; s < WSTAGES; s++) { mb_init(full + s, 1); mb_init(empty + s, NWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % WSTAGES;
    mb_expect(full + s, SLOT);
    tma2d(stage + (size_t)s * SLOT, &tmX, 2 * g0, t * C, full + s);
  };
  if (tid == 0) for (int t = 0; t < WSTAGES - 1 && t < T; t++) issue(t);

  // this lane's channels: c = l + L i
  float2 w[8], uw[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    w[i] = __ldg(a.W + (size_t)(l + L * i) * a.Gp + g);
    if (MODE == 0 && a.WL != nullptr && k != 0) { float2 wl = __ldg(a.WL + (size_t)(l + L * i) * a.Gp + g); w[i] = csub(w[i], wl); }
    uw[i] = make_float2(0.f, 0.f);
  }
  if (MODE == 0 && a.normalize_weight && k != 0 && a.kind != BTKB_BF_DS) {
    float nrm = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) nrm = fmaf(w[i].x, w[i].x, fmaf(w[i].y, w[i].y, nrm));
    nrm = red<L>(nrm, gm);
    const float sc = 1.0f / (sqrtf(nrm) * (float)a.Ctrue);   // (zero-padded channel rows do not count, see PerBinArgs::Ctrue)
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i].x *= sc; w[i].y *= sc; }
  }
  float se = a.lms.init_diagonal_load, Eavg = a.lms.init_diagonal_load, gamma = a.lms.gamma;
  int n_updates = 0, slow_cnt = a.lms.slowdown_after + 1;
  const float one_m_beta = 1.0f - a.lms.beta, inv_sil = 1.0f / a.lms.sil_thresh;
  float e_next = (MODE == 1 && T > 0) ? __ldg(a.E + u) : 0.f;

  for (int t = 0; t < T; t++) {
    const int s = t % WSTAGES;
    if (tid == 0 && t + WSTAGES - 1 < T) {
      if (t >= 1) mb_wait(empty + ((t - 1) % WSTAGES), (uint32_t)(((t - 1) / WSTAGES) & 1));
      issue(t + WSTAGES - 1);
    }
    mb_wait(full + s, (uint32_t)((t / WSTAGES) & 1));
    float2 x[8];
    {
      const unsigned char* base = stage + (size_t)s * SLOT;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int r = l + L * i;
        x[i] = *reinterpret_cast<const float2*>(base + r * 128 + ((((j >> 1) ^ (r & 7)) << 4) | ((j & 1) << 3)));
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mb_arrive(empty + s);

    const float energy = e_next;
    if (MODE == 1 && t + 1 < T) e_next = __ldg(a.E + (size_t)(t + 1) * a.U + u);
    float2 y = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++) { if constexpr (PK) y = f2_cmac_conj(y, x[i], w[i]); else macc(y, x[i], w[i]); }
    y = red2<L>(y, gm);
    const bool live = t < Tu;
    if (MODE == 1) {
      if (--slow_cnt == 0) { gamma *= 0.5f; slow_cnt = a.lms.slowdown_after; }
      const bool adapt = energy > (Eavg * inv_sil);
      float nx = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) nx = fmaf(x[i].x, x[i].x, fmaf(x[i].y, x[i].y, nx));
      nx = red<L>(nx, gm);
      float sub = (t > 0) ? fmaf(se, a.lms.beta, one_m_beta * nx) : nx;
      sub = fmaxf(sub, a.lms.energy_floor);
      if (adapt && live) {   // uniform over the L lanes of a chain
        float2 ux = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i++) { if constexpr (PK) ux = f2_cmac(ux, uw[i], x[i]); else mac(ux, uw[i], x[i]); }
        ux = red2<L>(ux, gm);
        const float2 epa = csub(y, ux);
        const float alphaK = gamma / sub;
        const float2 cy = make_float2((float)a.Ctrue * y.x, (float)a.Ctrue * y.y);
        const float2 ea = make_float2(epa.x * alphaK, epa.y * alphaK);
        const float keep = (a.lms.regularization_param > 0.f) ? 1.0f - alphaK * a.lms.regularization_param : 1.0f;
        float n2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if constexpr (PK) {
            const float2 q = f2_sub_cmul(x[i], cy, w[i]);
            const float2 un = f2_add_cmulc(f2_scale(uw[i], keep), ea, q);
            uw[i] = un;
            n2 = fmaf(un.x, un.x, fmaf(un.y, un.y, n2));
            continue;
          }
          const float qx = fmaf(-cy.x, w[i].x, fmaf(cy.y, w[i].y, x[i].x));
          const float qy = fmaf(-cy.x, w[i].y, fmaf(-cy.y, w[i].x, x[i].y));
          const float unx = fmaf(ea.y, qy, fmaf(ea.x, qx, keep * uw[i].x));
          const float uny = fmaf(-ea.x, qy, fmaf(ea.y, qx, keep * uw[i].y));
          uw[i] = make_float2(unx, uny);
          n2 = fmaf(unx, unx, fmaf(uny, uny, n2));
        }
        n2 = red<L>(n2, gm);
        if (n2 > a.lms.max_wa_l2norm) {
          const float cK = sqrtf(a.lms.max_wa_l2norm / n2);
#pragma unroll
          for (int i = 0; i < 8; i++) { if constexpr (PK) uw[i] = f2_scale(uw[i], cK); else { uw[i].x *= cK; uw[i].y *= cK; } }
        }
        se = sub;
        n_updates++;
      }
      if (t >= a.lms.min_frames) {
        float2 ux = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i++) { if constexpr (PK) ux = f2_cmac(ux, uw[i], x[i]); else mac(ux, uw[i], x[i]); }
        ux = red2<L>(ux, gm);
        y = csub(y, ux);
      }
      Eavg = fmaf(Eavg, a.lms.beta, one_m_beta * energy);
    }
    if (valid && l == 0) a.Y[(size_t)t * a.Gp + g] = live ? y : make_float2(0.f, 0.f);
  }
  if (MODE == 1 && valid) {
    if (a.UA != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; i++) a.UA[(size_t)(l + L * i) * a.Gp + g] = uw[i];
    }
    if (k == 0 && l == 0 && a.stats_updates != nullptr) a.stats_updates[u] = (float)n_updates;
  }
}

// One CTA (256 threads) per chain: R += x x^H over the frames flagged in noise_mask (pybeamformer.py:976-982), any C <= 64.
__global__ void __launch_bounds__(256) k_covariance_wide(PerBinArgs a) {
  __shared__ float2 xs[64];
  const int g = blockIdx.x;
  const int u = g / a.K, C = a.C;
  const int NE = C * C;                         // matrix entries; thread handles e = tid, tid+256, ...
  float2 acc[16];
#pragma unroll
  for (int q = 0; q < 16; q++) acc[q] = make_float2(0.f, 0.f);
  for (int t = 0; t < a.T; t++) {
    if (!a.noise_mask[(size_t)t * a.U + u]) continue;   // uniform over the CTA
    __syncthreads();
    if (threadIdx.x < C) xs[threadIdx.x] = a.X[((size_t)t * C + threadIdx.x) * a.Gp + g];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int e = threadIdx.x + q * 256;
      if (e < NE) macc(acc[q], xs[e / C], xs[e % C]);
    }
  }
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const int e = threadIdx.x + q * 256;
    if (e < NE) a.R[(size_t)e * a.Gp + g] = acc[q];
  }
}

// One CTA (C threads, one per row) per chain: w = (R^H)^-1 d / (C d^H R^-1 d) by Gaussian elimination with partial
// pivoting in double precision, matrix in shared memory (beamformer.cc:2350-2402; bin 0: all ones).
struct cdw { double x, y; };
__device__ __forceinline__ cdw cw(double x, double y) { cdw r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cdw cwmul(cdw a, cdw b) { return cw(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cdw cwsub(cdw a, cdw b) { return cw(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cdw cwmsub(cdw a, cdw f, cdw b) {   // a - f b as four FMAs
  return cw(fma(-f.x, b.x, fma(f.y, b.y, a.x)), fma(-f.x, b.y, fma(-f.y, b.x, a.y)));
}
__device__ __forceinline__ cdw cwdiv(cdw a, cdw b) { double d = b.x * b.x + b.y * b.y; return cw((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d); }

// blockDim.x = 4 C threads: thread (r = tid % C, q = tid / C) owns the columns c = q (mod 4) of row r, so a warp covers 32
// rows and the elimination of one column is spread over 4 C threads (8 warps at C = 64: enough to hide the fp64 / shared
// memory latency that one thread per row could not; 34 ms -> 23 ms -> see profiles/ at configs[3]).  Per column: warp arg-max
// pivot search on the q = 0 threads combined through shared memory, distributed row swap, elimination; two barriers per
// column.  The back substitution runs column-oriented with every row updating itself.
constexpr int SOLVE_Q = 4;
// `list` non-null: the CTAs walk the chains list[1 .. list[0]] that a Cholesky kernel flagged (a small fixed grid: an empty list costs
// nothing, where one CTA per chain — each claiming the 66 KB matrix buffer just to find its flag clear — cost 4.7 ms at configs[3] size).
__global__ void k_mvdr_solve_wide(const float2* R, const float2* Dm, float2* W, const int* noise_count, int U, int C, int K, int Gp, float mu, int normalize, const int* list) {
  extern __shared__ __align__(16) unsigned char sm[];
  cdw* A = reinterpret_cast<cdw*>(sm);          // [C][C+1] augmented, row-major
  __shared__ double wbest[2]; __shared__ int widx[2];
  __shared__ double lam_part[2][2];
  const int tid = threadIdx.x;
  const int r = tid % C, q = tid / C;           // blockDim.x == SOLVE_Q * C
  const int nwarp0 = (C + 31) / 32;             // warps holding the q = 0 threads
  const int nwork = list ? list[0] : U * K;
 for (int it = blockIdx.x; it < nwork; it += gridDim.x) {   // (every exit below is uniform over the CTA)
  __syncthreads();                              // the previous chain's buffers are free
  const int g = list ? list[1 + it] : it;
  const int u = g / K, k = g - u * K;
  if (k == 0) { if (tid < C) W[(size_t)tid * Gp + g] = make_float2(1.f, 0.f); continue; }
  double scale = 1.0;
  if (normalize && noise_count != nullptr && noise_count[u] > 0) scale = 1.0 / (double)noise_count[u];
  const int LD = C + 1;
  // A = R^H with loading; row r of A = conj of column r of R
  for (int c = q; c < C; c += SOLVE_Q) {
    float2 t = R[(size_t)(c * C + r) * Gp + g];
    cdw v = cw((double)t.x * scale, -(double)t.y * scale);
    if (c == r) v.x += (double)mu;
    A[r * LD + c] = v;
  }
  if (q == 0) { float2 t = Dm[(size_t)r * Gp + g]; A[r * LD + C] = cw(t.x, t.y); }
  __syncthreads();
  bool singular = false;
  for (int col = 0; col < C; col++) {
    // pivot: largest |A[row][col]|, row >= col (ties: the smallest row index, like a serial scan)
    if (tid < 32 * nwarp0) {
      double m2 = -1.0; int idx = tid;
      if (tid < C && tid >= col) { const cdw v = A[tid * LD + col]; m2 = v.x * v.x + v.y * v.y; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(0xffffffffu, m2, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > m2 || (om == m2 && oi < idx)) { m2 = om; idx = oi; }
      }
      if ((tid & 31) == 0) { wbest[tid >> 5] = m2; widx[tid >> 5] = idx; }
    }
    __syncthreads();
    double best = wbest[0]; int piv = widx[0];
    if (nwarp0 > 1 && (wbest[1] > best)) { best = wbest[1]; piv = widx[1]; }
    if (!(best > 1e-60)) { singular = true; break; }   // uniform over the CTA
    if (piv != col)
      for (int c = tid; c <= C; c += blockDim.x) { cdw t = A[col * LD + c]; A[col * LD + c] = A[piv * LD + c]; A[piv * LD + c] = t; }
    __syncthreads();
    if (r > col) {
      const cdw f = cwdiv(A[r * LD + col], A[col * LD + col]);   // column `col` of this row is not written in this step
#pragma unroll 4
      for (int c = col + 1 + q; c <= C; c += SOLVE_Q) A[r * LD + c] = cwmsub(A[r * LD + c], f, A[col * LD + c]);
    }
    __syncthreads();   // rows are shared by four threads now: the next pivot search reads what the other three wrote
  }
  __syncthreads();
  if (!singular) {
    // column-oriented back substitution into column C: x_i = b_i / a_ii, then every row r < i subtracts a_ri x_i from its b_r
    for (int i = C - 1; i >= 0; i--) {
      if (tid == i) A[i * LD + C] = cwdiv(A[i * LD + C], A[i * LD + i]);
      __syncthreads();
      if (tid < i) A[tid * LD + C] = cwmsub(A[tid * LD + C], A[tid * LD + i], A[i * LD + C]);
    }
  } else if (tid < C) {
    float2 t = Dm[(size_t)tid * Gp + g]; A[tid * LD + C] = cw(t.x, t.y);   // identity fallback (beamformer.cc:2381-2383)
  }
  // Lambda = t^H d (times C)
  if (tid < 32 * nwarp0) {
    double lr = 0.0, li = 0.0;
    if (tid < C) { const float2 d = Dm[(size_t)tid * Gp + g]; const cdw tv = A[tid * LD + C]; lr = tv.x * d.x + tv.y * d.y; li = tv.x * d.y - tv.y * d.x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lr += __shfl_xor_sync(0xffffffffu, lr, o); li += __shfl_xor_sync(0xffffffffu, li, o); }
    if ((tid & 31) == 0) { lam_part[tid >> 5][0] = lr; lam_part[tid >> 5][1] = li; }
  }
  __syncthreads();
  double lam_re = lam_part[0][0], lam_im = lam_part[0][1];
  if (nwarp0 > 1) { lam_re += lam_part[1][0]; lam_im += lam_part[1][1]; }
  if (tid < C) {
    const cdw wv = cwdiv(A[tid * LD + C], cw(lam_re * C, lam_im * C));
    W[(size_t)tid * Gp + g] = make_float2((float)wv.x, (float)wv.y);
  }
 }
}

// The same elimination with IMPLICIT pivoting and ONE barrier per column (k_mvdr_solve_wide above: three — pivot search, row swap,
// elimination).  Rows are never swapped: a row that has served as pivot is marked done and keeps its place, order[col] remembers which
// row eliminated column col; and the pivot search for column col + 1 rides on the elimination of column col (the q = 0 thread of a row
// has the row's new entry of column col + 1 in a register the moment it is computed), so the arg-max shuffle needs no barrier of its
// own and the single barrier publishes both the eliminated entries and the candidates.  Same pivots, same arithmetic per entry as the
// swapping version (partial pivoting by largest modulus, ties to the smallest row index).
__global__ void k_mvdr_solve_wide_ip(const float2* R, const float2* Dm, float2* W, const int* noise_count, int U, int C, int K, int Gp, float mu, int normalize, const unsigned char* todo) {
  extern __shared__ __align__(16) unsigned char sm[];
  if (todo != nullptr && !todo[blockIdx.x]) return;
  cdw* A = reinterpret_cast<cdw*>(sm);          // [C][C+1] augmented, row-major
  __shared__ double wbest[2][2]; __shared__ int widx[2][2];
  __shared__ int order[64];
  __shared__ double lam_part[2][2];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int r = tid % C, q = tid / C;           // blockDim.x == SOLVE_Q * C
  const int u = g / K, k = g - u * K;
  const int nwarp0 = (C + 31) / 32;
  if (k == 0) { if (tid < C) W[(size_t)tid * Gp + g] = make_float2(1.f, 0.f); return; }
  double scale = 1.0;
  if (normalize && noise_count != nullptr && noise_count[u] > 0) scale = 1.0 / (double)noise_count[u];
  const int LD = C + 1;
  double cand = -1.0;                           // |A[r][col]|^2 of the column about to be eliminated (q == 0 threads)
  for (int c = q; c < C; c += SOLVE_Q) {
    float2 t = R[(size_t)(c * C + r) * Gp + g];
    cdw v = cw((double)t.x * scale, -(double)t.y * scale);
    if (c == r) v.x += (double)mu;
    A[r * LD + c] = v;
    if (c == 0) cand = v.x * v.x + v.y * v.y;
  }
  if (q == 0) { float2 t = Dm[(size_t)r * Gp + g]; A[r * LD + C] = cw(t.x, t.y); }
  bool done = false, singular = false;
  int mypos = -1;
  for (int col = 0; col < C; col++) {
    if (tid < 32 * nwarp0) {
      double m2 = (tid < C && !done) ? cand : -1.0; int idx = tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(0xffffffffu, m2, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > m2 || (om == m2 && oi < idx)) { m2 = om; idx = oi; }
      }
      if ((tid & 31) == 0) { wbest[col & 1][tid >> 5] = m2; widx[col & 1][tid >> 5] = idx; }
    }
    __syncthreads();   // the one barrier of the column: eliminated entries of column col - 1 and the pivot candidates are visible
    double best = wbest[col & 1][0]; int piv = widx[col & 1][0];
    if (nwarp0 > 1 && (wbest[col & 1][1] > best)) { best = wbest[col & 1][1]; piv = widx[col & 1][1]; }
    if (!(best > 1e-60)) { singular = true; break; }   // uniform over the CTA
    if (r == piv) { done = true; mypos = col; }
    if (tid == 0) order[col] = piv;
    if (!done) {
      const cdw f = cwdiv(A[r * LD + col], A[piv * LD + col]);   // neither entry is written in this step
#pragma unroll 4
      for (int c = col + 1 + q; c <= C; c += SOLVE_Q) {
        const cdw v = cwmsub(A[r * LD + c], f, A[piv * LD + c]);
        A[r * LD + c] = v;
        if (c == col + 1) cand = v.x * v.x + v.y * v.y;           // (q == 0 only) the candidate for the next column
      }
    }
  }
  __syncthreads();
  if (!singular) {
    // back substitution in pivot order: x_col = b[p] / a[p][col] with p = order[col]; every earlier pivot row subtracts a[.][col] x_col
    for (int col = C - 1; col >= 0; col--) {
      const int p = order[col];
      if (q == 0 && r == p) A[p * LD + C] = cwdiv(A[p * LD + C], A[p * LD + col]);
      __syncthreads();
      if (q == 0 && mypos < col) A[r * LD + C] = cwmsub(A[r * LD + C], A[r * LD + col], A[p * LD + C]);
    }
    __syncthreads();
  }
  // unknown c sits in the right-hand side of its pivot row (identity fallback: t = d, beamformer.cc:2381-2383)
  cdw tv = cw(0.0, 0.0);
  if (tid < C) {
    if (singular) { const float2 t = Dm[(size_t)tid * Gp + g]; tv = cw(t.x, t.y); }
    else tv = A[order[tid] * LD + C];
  }
  if (tid < 32 * nwarp0) {
    double lr = 0.0, li = 0.0;
    if (tid < C) { const float2 d = Dm[(size_t)tid * Gp + g]; lr = tv.x * d.x + tv.y * d.y; li = tv.x * d.y - tv.y * d.x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lr += __shfl_xor_sync(0xffffffffu, lr, o); li += __shfl_xor_sync(0xffffffffu, li, o); }
    if ((tid & 31) == 0) { lam_part[tid >> 5][0] = lr; lam_part[tid >> 5][1] = li; }
  }
  __syncthreads();
  double lam_re = lam_part[0][0], lam_im = lam_part[0][1];
  if (nwarp0 > 1) { lam_re += lam_part[1][0]; lam_im += lam_part[1][1]; }
  if (tid < C) {
    const cdw wv = cwdiv(tv, cw(lam_re * C, lam_im * C));
    W[(size_t)tid * Gp + g] = make_float2((float)wv.x, (float)wv.y);
  }
}

// Hermitian positive-definite case of the same solve — every matrix the path itself builds (sample covariance, diffuse model, loaded
// versions of them) — as a Cholesky factorisation with ONE WARP PER CHAIN: lower triangle packed in shared memory (C (C + 1) / 2
// complex doubles, 33 KiB at C = 64 instead of the 66 KiB augmented square), lane l owns rows l and C - 1 - l (balanced: C - 1 - 2 j
// trailing entries per column for every lane), dependent steps separated by __syncwarp instead of CTA barriers (the LU kernel above
// spends its time in 4 x 64 of those), half the flops, no pivot search.  A chain whose matrix is not Hermitian (|a_ij - conj(a_ji)|^2
// > 1e-10 a_ii a_jj) or not positive definite is appended to `list` and left to the LU kernel, which then runs only for those chains.
// R^H t = d with R Hermitian is R t = d;  w = t / (C t^H d)  (beamformer.cc:2386-2398), bin 0: all ones.
constexpr int CHOL_WARPS = 2;
__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i
__global__ void __launch_bounds__(32 * CHOL_WARPS) k_mvdr_solve_wide_chol(const float2* R, const float2* Dm, float2* W, const int* noise_count, int U, int C, int K, int Gp,
                                                                          float mu, int normalize, int* list) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * CHOL_WARPS + warp;
  if (g >= U * K) return;
  const int NT = C * (C + 1) / 2;
  cdw* Lm = reinterpret_cast<cdw*>(sm) + (size_t)warp * (NT + C);   // packed lower triangle, then the right-hand side / solution [C]
  cdw* bv = Lm + NT;
  const int u = g / K, k = g - u * K;
  if (k == 0) { for (int c = lane; c < C; c += 32) W[(size_t)c * Gp + g] = make_float2(1.f, 0.f); return; }
  double scale = 1.0;
  if (normalize && noise_count != nullptr && noise_count[u] > 0) scale = 1.0 / (double)noise_count[u];
  // rows of this lane: r0 = lane, r1 = C - 1 - lane (C = 16: lanes >= 8 own nothing; C = 32: lanes >= 16 own nothing)
  const int nrow = (2 * lane < C) ? ((C - 1 - lane != lane) ? 2 : 1) : 0;
  const int rows[2] = {lane, C - 1 - lane};
  bool herm = true;
  for (int q = 0; q < nrow; q++) {
    const int i = rows[q];
    for (int j = 0; j <= i; j++) {
      const float2 a = R[(size_t)(i * C + j) * Gp + g], b = R[(size_t)(j * C + i) * Gp + g];
      cdw v = cw(0.5 * ((double)a.x + (double)b.x) * scale, 0.5 * ((double)a.y - (double)b.y) * scale);   // Hermitian part
      if (i == j) { v.x += (double)mu; v.y = 0.0; }
      Lm[tri(i, j)] = v;
      const double dr = (double)a.x - (double)b.x, di = (double)a.y + (double)b.y;
      const double aii = (double)R[(size_t)(i * C + i) * Gp + g].x, ajj = (double)R[(size_t)(j * C + j) * Gp + g].x;
      if (dr * dr + di * di > 1e-10 * fabs(aii * ajj) + 1e-300) herm = false;
    }
    const float2 t = Dm[(size_t)i * Gp + g];
    bv[i] = cw(t.x, t.y);
  }
  bool ok = __all_sync(0xffffffffu, herm);
  __syncwarp();
  // ---- Cholesky, right-looking, column by column
  for (int j = 0; j < C && ok; j++) {
    const double d = Lm[tri(j, j)].x;           // broadcast read
    if (!(d > 0.0)) { ok = false; break; }      // uniform over the warp
    const double sj = sqrt(d), inv = 1.0 / sj;
    __syncwarp();                                // everybody has read the pivot
    for (int q = 0; q < nrow; q++) {
      const int i = rows[q];
      if (i == j) Lm[tri(j, j)] = cw(sj, 0.0);
      else if (i > j) { cdw v = Lm[tri(i, j)]; Lm[tri(i, j)] = cw(v.x * inv, v.y * inv); }
    }
    __syncwarp();
    for (int q = 0; q < nrow; q++) {
      const int i = rows[q];
      if (i <= j) continue;
      const cdw lij = Lm[tri(i, j)];
      cdw* row = Lm + tri(i, 0);
      // a_ic -= l_ij conj(l_cj), c = j+1 .. i; four entries per step with all loads issued before the first store (the compiler cannot
      // prove that row[] and the column entries do not alias, and one dependent load -> FMA -> store chain per entry left the warp
      // waiting on shared-memory latency)
      int c = j + 1;
      for (; c + 3 <= i; c += 4) {
        const cdw l0 = Lm[tri(c, j)], l1 = Lm[tri(c + 1, j)], l2 = Lm[tri(c + 2, j)], l3 = Lm[tri(c + 3, j)];
        cdw v0 = row[c], v1 = row[c + 1], v2 = row[c + 2], v3 = row[c + 3];
        v0.x = fma(-lij.x, l0.x, fma(-lij.y, l0.y, v0.x)); v0.y = fma(-lij.y, l0.x, fma(lij.x, l0.y, v0.y));
        v1.x = fma(-lij.x, l1.x, fma(-lij.y, l1.y, v1.x)); v1.y = fma(-lij.y, l1.x, fma(lij.x, l1.y, v1.y));
        v2.x = fma(-lij.x, l2.x, fma(-lij.y, l2.y, v2.x)); v2.y = fma(-lij.y, l2.x, fma(lij.x, l2.y, v2.y));
        v3.x = fma(-lij.x, l3.x, fma(-lij.y, l3.y, v3.x)); v3.y = fma(-lij.y, l3.x, fma(lij.x, l3.y, v3.y));
        if (c + 3 == i) v3.y = 0.0;
        row[c] = v0; row[c + 1] = v1; row[c + 2] = v2; row[c + 3] = v3;
      }
      for (; c <= i; c++) {
        const cdw lcj = Lm[tri(c, j)];
        cdw v = row[c];
        v.x = fma(-lij.x, lcj.x, fma(-lij.y, lcj.y, v.x));
        v.y = fma(-lij.y, lcj.x, fma(lij.x, lcj.y, v.y));
        if (c == i) v.y = 0.0;
        row[c] = v;
      }
    }
    __syncwarp();
  }
  if (!ok) { if (lane == 0) list[1 + atomicAdd(list, 1)] = g; return; }   // left to the pivoted LU
  // ---- L y = d (forward), L^H t = y (backward); column-oriented, every lane updates its own rows
  for (int i = 0; i < C; i++) {
    const cdw yi = cw(bv[i].x / Lm[tri(i, i)].x, bv[i].y / Lm[tri(i, i)].x);   // every lane computes the same value from broadcast reads
    __syncwarp();
    for (int q = 0; q < nrow; q++) {
      const int r = rows[q];
      if (r == i) bv[i] = yi;
      else if (r > i) bv[r] = cwmsub(bv[r], Lm[tri(r, i)], yi);
    }
    __syncwarp();
  }
  for (int i = C - 1; i >= 0; i--) {
    const cdw ti = cw(bv[i].x / Lm[tri(i, i)].x, bv[i].y / Lm[tri(i, i)].x);
    __syncwarp();
    for (int q = 0; q < nrow; q++) {
      const int r = rows[q];
      if (r == i) bv[i] = ti;
      else if (r < i) { const cdw l = Lm[tri(i, r)]; bv[r] = cwmsub(bv[r], cw(l.x, -l.y), ti); }   // y_r -= conj(l_ir) t_i
    }
    __syncwarp();
  }
  // ---- Lambda = t^H d, w = t / (C Lambda)
  double lr = 0.0, li = 0.0;
  for (int q = 0; q < nrow; q++) {
    const int r = rows[q];
    const float2 dd = Dm[(size_t)r * Gp + g]; const cdw tv = bv[r];
    lr += tv.x * dd.x + tv.y * dd.y; li += tv.x * dd.y - tv.y * dd.x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lr += __shfl_xor_sync(0xffffffffu, lr, o); li += __shfl_xor_sync(0xffffffffu, li, o); }
  for (int q = 0; q < nrow; q++) {
    const int r = rows[q];
    const cdw wv = cwdiv(bv[r], cw(lr * C, li * C));
    W[(size_t)r * Gp + g] = make_float2((float)wv.x, (float)wv.y);
  }
}

// The same Hermitian positive-definite solve as a BLOCKED Cholesky on the fp64 tensor cores, one CTA of four warps per chain, the whole
// lower block triangle in shared memory (C = 16 NB: NB (NB + 1) / 2 blocks of 16 x 16, two planes (re, im), row stride 20 doubles,
// column XOR-swizzled with the row inside the block: the layout of btkb_wpe.cu's panel, conflict-free both for row-per-lane sweeps and
// for mma fragments; 51 KB at C = 64, four CTAs per SM).  Per 16-column panel: the diagonal block by one warp (column sweep, pairs of the
// rank-one update dealt out to the lanes), L21 = A21 L11^-H one thread per row in registers, A22 -= L21 L21^H as mma.sync m8n8k4 f64 on
// 16 x 16 blocks — nothing leaves the SM between loading R and storing w.  Then L y = d and L^H t = y panel by panel (the 16 x 16 solves
// by shuffles in one warp, the rest one thread per row), w = t / (C t^H d).  Chains that are not Hermitian positive definite are flagged
// for the pivoted LU exactly like k_mvdr_solve_wide_chol does.
constexpr int BLK_THREADS = 128;
constexpr int BLK_PS = 20;
__device__ __forceinline__ int blk_idx(int r, int j) {   // block column of j <= block row of r
  const int bi = r >> 4, bk = j >> 4;
  return (bi * (bi + 1) / 2 + bk) * (16 * BLK_PS) + (r & 15) * BLK_PS + ((j & 15) ^ (r & 15));
}
__device__ __forceinline__ void dmma884w(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(BLK_THREADS, 4) k_mvdr_solve_wide_blk(const float2* R, const float2* Dm, float2* W, const int* noise_count, int U, int C, int K, int Gp,
                                                                        float mu, int normalize, int* list) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ int s_ok, s_herm;   // pivots positive so far / matrix Hermitian (two flags: each is written in one phase and read after the barrier that ends it)
  __shared__ double s_lam[2][BLK_THREADS / 32];
  __shared__ unsigned char s_tri[120][2];   // (row, column) of the e-th entry of a 15 x 15 lower triangle: the pairs of the rank-one updates
  const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NB = C >> 4, plane = NB * (NB + 1) / 2 * 16 * BLK_PS;
  double* Mre = reinterpret_cast<double*>(sm);
  double* Mim = Mre + plane;
  cdw* bv = reinterpret_cast<cdw*>(Mim + plane);   // [C] right-hand side / solution
  double* dg = reinterpret_cast<double*>(bv + C);  // [C] the raw diagonal (Hermitian test)
  double* dinv = dg + C;                           // [16] 1 / L_jj of the current diagonal block
  auto ld = [&](int r, int j) -> cdw { const int i = blk_idx(r, j); return cw(Mre[i], Mim[i]); };
  auto st = [&](int r, int j, cdw v) { const int i = blk_idx(r, j); Mre[i] = v.x; Mim[i] = v.y; };
  const int u = g / K, k = g - u * K;
  if (tid == 0) { s_ok = 1; s_herm = 1; }
  if (k == 0) { for (int c = tid; c < C; c += BLK_THREADS) W[(size_t)c * Gp + g] = make_float2(1.f, 0.f); return; }
  if (tid < 120) {
    int rr = 0;
    while ((rr + 1) * (rr + 2) / 2 <= tid) rr++;
    s_tri[tid][0] = (unsigned char)rr; s_tri[tid][1] = (unsigned char)(tid - rr * (rr + 1) / 2);
  }
  double scale = 1.0;
  if (normalize && noise_count != nullptr && noise_count[u] > 0) scale = 1.0 / (double)noise_count[u];
  for (int c = tid; c < C; c += BLK_THREADS) {
    dg[c] = (double)R[(size_t)(c * C + c) * Gp + g].x;
    const float2 t = Dm[(size_t)c * Gp + g];
    bv[c] = cw(t.x, -t.y);   // the right-hand side rides along as the row vector a = d^H (x L^H = a  <=>  x = (L^-1 d)^H): forward solve for free
  }
  __syncthreads();
  // ---- load the Hermitian part of the lower triangle (+ mu on the diagonal); four pairs of loads in flight per thread
  bool herm = true;
  for (int e0 = tid; e0 < C * C; e0 += 4 * BLK_THREADS) {
    float2 a4[4], b4[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int e = e0 + q * BLK_THREADS, i = e / C, j = e - i * C;
      if (e < C * C && j <= i) { a4[q] = R[(size_t)(i * C + j) * Gp + g]; b4[q] = R[(size_t)(j * C + i) * Gp + g]; }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int e = e0 + q * BLK_THREADS, i = e / C, j = e - i * C;
      if (e < C * C && j <= i) {
        cdw v = cw(0.5 * ((double)a4[q].x + (double)b4[q].x) * scale, 0.5 * ((double)a4[q].y - (double)b4[q].y) * scale);
        if (i == j) { v.x += (double)mu; v.y = 0.0; }
        st(i, j, v);
        const double dr = (double)a4[q].x - (double)b4[q].x, di = (double)a4[q].y + (double)b4[q].y;
        if (dr * dr + di * di > 1e-10 * fabs(dg[i] * dg[j]) + 1e-300) herm = false;
      }
    }
  }
  if (!herm) s_herm = 0;
  __syncthreads();
  if (!s_herm) { if (tid == 0) list[1 + atomicAdd(list, 1)] = g; return; }   // left to the pivoted LU

  // ---- blocked right-looking Cholesky.  Per panel: [diagonal block, one warp] -> TRSM of the rows below (and of the right-hand-side row)
  // -> trailing update.  From the second panel on the diagonal block is factored AHEAD: warp 0 takes it out of the trailing update of the
  // previous panel first and factors it while warps 1 .. 3 update the other blocks and the right-hand-side row.
  auto factor_diag = [&](int j0) {   // warp 0, all 32 lanes
    for (int jj = 0; jj < 16; jj++) {
      const double dj = Mre[blk_idx(j0 + jj, j0 + jj)];
      if (!(dj > 0.0)) s_ok = 0;
      const double inv = rsqrt(fmax(dj, 1e-300));
      __syncwarp();   // every lane has read the pivot before its owner overwrites it
      if (lane >= jj && lane < 16) {
        if (lane == jj) { st(j0 + lane, j0 + jj, cw(dj * inv, 0.0)); dinv[jj] = inv; }
        else { const cdw v = ld(j0 + lane, j0 + jj); st(j0 + lane, j0 + jj, cw(v.x * inv, v.y * inv)); }
      }
      __syncwarp();
      const int m = 15 - jj;
      for (int e = lane; e < m * (m + 1) / 2; e += 32) {
        const int r = j0 + jj + 1 + s_tri[e][0], kk = j0 + jj + 1 + s_tri[e][1];
        const cdw lr_ = ld(r, j0 + jj), lk = ld(kk, j0 + jj);
        cdw v = ld(r, kk);
        v.x = fma(-lr_.x, lk.x, fma(-lr_.y, lk.y, v.x)); v.y = fma(-lr_.y, lk.x, fma(lr_.x, lk.y, v.y));   // v -= l_r conj(l_k)
        if (kk == r) v.y = 0.0;
        st(r, kk, v);
      }
      __syncwarp();
    }
  };
  if (warp == 0) factor_diag(0);
  __syncthreads();
  for (int p = 0; p < NB; p++) {
    const int j0 = 16 * p;
    if (!s_ok) { if (tid == 0) list[1 + atomicAdd(list, 1)] = g; return; }   // left to the pivoted LU (uniform: s_ok was last written before a barrier)
    // L21 = A21 L11^-H: one thread per row below the block, the row's 16 entries in registers, right-looking; the thread after the
    // last row does the same for the right-hand-side row (x_p = a_p L11^-H)
    {
      const int r = j0 + 16 + tid;
      if (r <= C) {
        const bool rhs = r == C;
        cdw v[16];
#pragma unroll
        for (int jj = 0; jj < 16; jj++) v[jj] = rhs ? bv[j0 + jj] : ld(r, j0 + jj);
#pragma unroll
        for (int jj = 0; jj < 16; jj++) {
          const double inv = dinv[jj];
          v[jj] = cw(v[jj].x * inv, v[jj].y * inv);
#pragma unroll
          for (int q = 0; q < 16; q++) {
            if (q > jj) {
              const cdw l = ld(j0 + q, j0 + jj);
              v[q].x = fma(-v[jj].x, l.x, fma(-v[jj].y, l.y, v[q].x)); v[q].y = fma(-v[jj].y, l.x, fma(v[jj].x, l.y, v[q].y));   // v_q -= x_jj conj(L[q][jj])
            }
          }
          if (rhs) bv[j0 + jj] = v[jj]; else st(r, j0 + jj, v[jj]);
        }
      }
    }
    __syncthreads();
    // A22 -= L21 L21^H on 16 x 16 blocks (bi >= bk > p).  Block 0 = the next diagonal block: warp 0 updates it and factors it right away;
    // the other blocks go round warps 1 .. 3, which also take the right-hand-side row: a_c -= sum_jj x_p[jj] conj(L[c][j0 + jj]), c >= j0 + 16
    {
      const int nt = NB - 1 - p, lr = lane >> 2, lc = lane & 3;
      const int npair = nt * (nt + 1) / 2;
      for (int bp = (warp == 0) ? 0 : warp; bp < npair; bp += (warp == 0) ? npair : BLK_THREADS / 32 - 1) {
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= bp) bi++;
        const int bk = bp - bi * (bi + 1) / 2;
        const int ri = (p + 1 + bi) * 16, rk = (p + 1 + bk) * 16;
        double sr[2][2][2], si[2][2][2];
#pragma unroll
        for (int ti = 0; ti < 2; ti++)
#pragma unroll
          for (int tk = 0; tk < 2; tk++) { sr[ti][tk][0] = sr[ti][tk][1] = 0.0; si[ti][tk][0] = si[ti][tk][1] = 0.0; }
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += 4) {
          double ar[2], ai[2], br[2], bim[2], nbi[2];
#pragma unroll
          for (int t2 = 0; t2 < 2; t2++) {
            const int ia = blk_idx(ri + t2 * 8 + lr, j0 + k0 + lc), ib = blk_idx(rk + t2 * 8 + lr, j0 + k0 + lc);
            ar[t2] = Mre[ia]; ai[t2] = Mim[ia]; br[t2] = Mre[ib]; bim[t2] = Mim[ib]; nbi[t2] = -bim[t2];
          }
#pragma unroll
          for (int ti = 0; ti < 2; ti++)
#pragma unroll
            for (int tk = 0; tk < 2; tk++) {
              dmma884w(sr[ti][tk][0], sr[ti][tk][1], ar[ti], br[tk]);
              dmma884w(sr[ti][tk][0], sr[ti][tk][1], ai[ti], bim[tk]);
              dmma884w(si[ti][tk][0], si[ti][tk][1], ai[ti], br[tk]);
              dmma884w(si[ti][tk][0], si[ti][tk][1], ar[ti], nbi[tk]);
            }
        }
#pragma unroll
        for (int ti = 0; ti < 2; ti++)
#pragma unroll
          for (int tk = 0; tk < 2; tk++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int row = ri + ti * 8 + lr, col = rk + tk * 8 + 2 * lc + e;
              if (col <= row) {
                const int i = blk_idx(row, col);
                Mre[i] -= sr[ti][tk][e]; Mim[i] = (col == row) ? 0.0 : Mim[i] - si[ti][tk][e];
              }
            }
      }
      if (warp == 0) { if (nt > 0) { __syncwarp(); factor_diag(j0 + 16); } }
      else {
        for (int c = j0 + 16 + (tid - 32); c < C; c += BLK_THREADS - 32) {
          cdw acc = bv[c];
#pragma unroll
          for (int jj = 0; jj < 16; jj++) { const cdw l = ld(c, j0 + jj); acc = cwmsub(acc, bv[j0 + jj], cw(l.x, -l.y)); }   // a_c -= x_jj conj(L[c][j0 + jj])
          bv[c] = acc;
        }
      }
    }
    __syncthreads();
  }
  // y = L^-1 d = conj(x)
  for (int c = tid; c < C; c += BLK_THREADS) bv[c].y = -bv[c].y;
  __syncthreads();

  // ---- L^H t = y, last panel first
  for (int p = NB - 1; p >= 0; p--) {
    const int j0 = 16 * p;
    if (warp == 0) {
      const int q = lane & 15;
      cdw y = (lane < 16) ? bv[j0 + q] : cw(0.0, 0.0);
      const double invq = 1.0 / Mre[blk_idx(j0 + q, j0 + q)];
      for (int jj = 15; jj >= 0; jj--) {
        const cdw tj = cw(__shfl_sync(0xffffffffu, y.x * invq, jj), __shfl_sync(0xffffffffu, y.y * invq, jj));
        if (q == jj) y = tj;
        else if (q < jj) { const cdw l = ld(j0 + jj, j0 + q); y = cwmsub(y, cw(l.x, -l.y), tj); }   // y_q -= conj(L[jj][q]) t_jj
      }
      if (lane < 16) bv[j0 + q] = y;
    }
    __syncthreads();
    {
      const int r = tid;
      if (r < j0) {
        cdw s = bv[r];
#pragma unroll
        for (int jj = 0; jj < 16; jj++) { const cdw l = ld(j0 + jj, r); s = cwmsub(s, cw(l.x, -l.y), bv[j0 + jj]); }
        bv[r] = s;
      }
    }
    __syncthreads();
  }
  // ---- Lambda = t^H d, w = t / (C Lambda)
  double lr = 0.0, li = 0.0;
  for (int r = tid; r < C; r += BLK_THREADS) {
    const float2 dd = Dm[(size_t)r * Gp + g]; const cdw tv = bv[r];
    lr += tv.x * dd.x + tv.y * dd.y; li += tv.x * dd.y - tv.y * dd.x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lr += __shfl_xor_sync(0xffffffffu, lr, o); li += __shfl_xor_sync(0xffffffffu, li, o); }
  if (lane == 0) { s_lam[0][warp] = lr; s_lam[1][warp] = li; }
  __syncthreads();
  lr = 0.0; li = 0.0;
  for (int w = 0; w < BLK_THREADS / 32; w++) { lr += s_lam[0][w]; li += s_lam[1][w]; }
  for (int r = tid; r < C; r += BLK_THREADS) {
    const cdw wv = cwdiv(bv[r], cw(lr * C, li * C));
    W[(size_t)r * Gp + g] = make_float2((float)wv.x, (float)wv.y);
  }
}

static cudaError_t make_map_wide(CUtensorMap* tm, const PerBinArgs& a, int C) {
  return encode_tensor_map_2d_f32(tm, a.X, (cuuint64_t)2 * a.Gp, (cuuint64_t)a.T * C, (cuuint64_t)a.Gp * sizeof(float2), (cuuint32_t)(2 * TC), (cuuint32_t)C,
                                  CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int L>
static cudaError_t launch_wide_l(const PerBinArgs& a, cudaStream_t st) {
  constexpr int C = 8 * L;
  const size_t smem = (size_t)C * TC * sizeof(float2) * WSTAGES + sizeof(uint64_t) * 2 * WSTAGES + 1024;
  CUtensorMap tm;
  cudaError_t e = make_map_wide(&tm, a, C);
  if (e != cudaSuccess) return e;
  const int grid = (a.G + TC - 1) / TC;
  const bool pk = env_packed("BTKB_PERBIN_PACKED");
  if (pk) {
    auto kern = (a.kind == BTKB_BF_GSC_LMS) ? k_perbin_wide<L, 1, true> : k_perbin_wide<L, 0, true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TC * L, smem, st>>>(tm, a);
  } else if (a.kind == BTKB_BF_GSC_LMS) {
    auto kern = k_perbin_wide<L, 1>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TC * L, smem, st>>>(tm, a);
  } else {
    auto kern = k_perbin_wide<L, 0>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TC * L, smem, st>>>(tm, a);
  }
  return cudaGetLastError();
}

}  // namespace wide

cudaError_t launch_perbin_wide(const PerBinArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  if (a.pf_kind != BTKB_PF_NONE) return cudaErrorInvalidValue;   // the CSD state of C > 8 channels does not fit registers
  switch (a.C) {
    case 16: return wide::launch_wide_l<2>(a, st);
    case 32: return wide::launch_wide_l<4>(a, st);
    case 64: return wide::launch_wide_l<8>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_covariance_wide(const PerBinArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.G <= 0) return cudaSuccess;
  if (a.C > 64) return cudaErrorInvalidValue;
  wide::k_covariance_wide<<<a.G, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_mvdr_solve_wide(const float2* R, const float2* D, float2* W, const int* noise_count, int U, int C, int K, int Gp, float mu,
                                   int normalize_by_count, int mode, int* list, cudaStream_t st) {
  // mode 0: pivoted LU for every chain (time independent of the matrices: 17.7 ms at configs[3] size).
  // mode 1 / 2: Hermitian positive-definite chains by a Cholesky kernel — 1: one warp per chain (16.9 ms), 2: blocked, on the fp64 tensor cores,
  //   matrix resident in shared memory (7.8 ms) — which appends every other chain to `list` (list[0] = count) for the LU, run by a small fixed
  //   grid over that list.  A batch of numerically indefinite matrices (62 noise frames for 64 channels with a 1e-4 loading) costs both passes;
  //   btkb_api.cu picks the mode per call from what it knows about the matrices (frames accumulated per utterance), BTKB_SOLVE_CHOL overrides.
  // mode 3: the LU with implicit pivoting (one barrier per column; 18.1 ms, profiles/r02q_*: the barrier COUNT is not what limits the LU —
  //   137 k warp instructions per matrix for 11 k warp-DFMAs of elimination work, issue slots 43 %, fp64 pipe 23 %).
  const size_t smem = sizeof(wide::cdw) * (size_t)C * (C + 1);
  cudaError_t e = cudaFuncSetAttribute(wide::k_mvdr_solve_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (mode == 3) {
    if ((e = cudaFuncSetAttribute(wide::k_mvdr_solve_wide_ip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    wide::k_mvdr_solve_wide_ip<<<U * K, wide::SOLVE_Q * C, smem, st>>>(R, D, W, noise_count, U, C, K, Gp, mu, normalize_by_count, nullptr);
    return cudaGetLastError();
  }
  if ((mode == 1 || mode == 2) && list != nullptr) {
    if ((e = cudaMemsetAsync(list, 0, sizeof(int), st)) != cudaSuccess) return e;
    if (mode == 2 && C % 16 == 0 && C <= 64) {
      const int NBk = C / 16;
      const size_t smb = (size_t)2 * (NBk * (NBk + 1) / 2) * 16 * wide::BLK_PS * sizeof(double) + (size_t)C * sizeof(wide::cdw) + (size_t)(C + 16) * sizeof(double);
      if ((e = cudaFuncSetAttribute(wide::k_mvdr_solve_wide_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb)) != cudaSuccess) return e;
      wide::k_mvdr_solve_wide_blk<<<U * K, wide::BLK_THREADS, smb, st>>>(R, D, W, noise_count, U, C, K, Gp, mu, normalize_by_count, list);
    } else {
      const size_t smc = sizeof(wide::cdw) * (size_t)wide::CHOL_WARPS * ((size_t)C * (C + 1) / 2 + C);
      if ((e = cudaFuncSetAttribute(wide::k_mvdr_solve_wide_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc)) != cudaSuccess) return e;
      wide::k_mvdr_solve_wide_chol<<<(U * K + wide::CHOL_WARPS - 1) / wide::CHOL_WARPS, 32 * wide::CHOL_WARPS, smc, st>>>(R, D, W, noise_count, U, C, K, Gp, mu, normalize_by_count, list);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int grid = std::min(U * K, 148 * 3);   // three 66 KB CTAs per SM: the whole list in flight at the LU's own occupancy
    wide::k_mvdr_solve_wide<<<grid, wide::SOLVE_Q * C, smem, st>>>(R, D, W, noise_count, U, C, K, Gp, mu, normalize_by_count, list);
    return cudaGetLastError();
  }
  wide::k_mvdr_solve_wide<<<U * K, wide::SOLVE_Q * C, smem, st>>>(R, D, W, noise_count, U, C, K, Gp, mu, normalize_by_count, nullptr);
  return cudaGetLastError();
}

}  // namespace btkb
