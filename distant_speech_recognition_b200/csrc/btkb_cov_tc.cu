// btkb_cov_tc.cu — K2w: 64-microphone spatial covariance on the 5th-generation tensor cores (tcgen05 + TMEM, sm_100a).
//
// Replaces, for C = 64, the accumulation loop of SubbandSMIMVDRBeamformer.accu_stats_from_label
// (btk20_src/lib/pybeamformer.py:948-992: R[m] += outer(x, conj(x)) per noise frame and bin) — at 64 microphones a genuine
// batched dense contraction (8 C^2 flop per bin-frame, AI ~ 60 flop/B on the snapshot tensor): one 64 x T x 64 complex
// Gram per (utterance, bin) chain, 65 792 chains at configs[3].
//
// Formulation.  With A_g = [Xr ; Xi] (128 real rows, K = frames) the real Gram S = A A^T holds every product:
//   Re R[c][c'] = S[Xr c][Xr c'] + S[Xi c][Xi c'],   Im R[c][c'] = S[Xi c][Xr c'] - S[Xr c][Xi c'].
// One tcgen05.mma (M = 128, N = 128, K = 8, kind::tf32, fp32 accumulate in TMEM) per 8 frames; A and B are THE SAME
// shared-memory tile (K-major, 128-byte swizzle).  TF32 keeps 10 mantissa bits, which is not enough for the 1e-4 parity budget
// after the MVDR solve, so every snapshot is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi,
// exact) and three MMAs accumulate hi hi^T + lo hi^T + hi lo^T (error ~ 2^-21 per product, fp32 class).
//
// Rows are interleaved in blocks of 8 ([Xr 0-7][Xi 0-7][Xr 8-15]...), so the TMEM lanes holding the Xr and the Xi row of a
// channel sit 8 lanes apart in the same warp and the epilogue combines them with warp shuffles (no shared-memory pass).
//
// CTA = 22 warps, persistent over groups of 2 adjacent chains:
//   (pre-pass)  k_cov_gather: series-major copy S[g][c][Ts] of X[t][c][g] through 32 x 32 shared-memory transposes.  A chain pair
//               is a 16-byte column of X; reading such a column directly costs 26 ms at configs[3] whether per-lane loads
//               (fully divergent) or a 3-D TMA box with a 16-byte inner extent do it (both measured), the coalesced pre-pass
//               plus 256-byte TMA rows does not.
//   warp 21     one thread loads the pair's K-block with two 2-D tensor-map TMAs: box {32 frames x (re, im), 64 channels} of S
//   warps 5-20  transposers: one LDS.64 per (chain, channel, lane = frame), noise-frame masking
//               (pybeamformer.py:963-975 via k_noise_mask), hi/lo split, st.shared into the swizzled operand tiles
//               (conflict-free: a warp writes one 128-byte row), fence.proxy.async + mbarrier arrive
//   warp 4      one elected thread issues the MMAs and tcgen05.commit's to the `empty` / `accumulator full` mbarriers
//   warps 0-3   epilogue: tcgen05.ld (32 lanes x 32 columns), shuffle-combine, R[(c C + c')][g] complex64 stores;
//               TMEM is double-buffered (2 x 2 chains x 128 columns = 512 columns), so it overlaps the next group's MMAs
// Shared memory: 4 operand stages x (hi + lo) x 16 KiB + 6 raw stages x 16 KiB = 224 KiB (the pipeline unit is one chain's K-block).
#include "btkb_internal.h"
#include "btkb_tensor_map.h"
#include <stdint.h>
#include <cstdlib>

namespace btkb {
namespace tc {

constexpr int C64 = 64;
constexpr int KB = 32;                     // frames per K-block = tf32 elements of one 128-byte operand row
constexpr int TILE_B = 128 * 128;          // one operand tile: 128 rows x 128 bytes
constexpr int NCH = 2;                     // chains per group
constexpr int STAGE_B = 2 * TILE_B;        // one pipeline unit = one chain's K-block: hi tile, lo tile
constexpr int NSTAGE = 4;                  // operand stages
constexpr int RAW_B = C64 * KB * 8;        // raw K-block of one chain: [c][t] complex64
constexpr int NRAW = 6;                    // raw stages
constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int TMA_WARP = EPI_WARPS + 1 + PROD_WARPS;
constexpr int THREADS = 32 * (EPI_WARPS + 1 + PROD_WARPS + 1);
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)NSTAGE * STAGE_B + (size_t)NRAW * RAW_B + 1024 /* alignment slack */ + 256 /* barriers */;

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
               "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mb_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
// 2-D tiled TMA load: box {2 KB floats (KB frames x (re, im)), 64 rows (channels)} of S viewed as float32 [G C][2 Ts] -> smem [c][t] complex64
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x0, int row0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(s_u32(dst)), "l"(tm), "r"(x0), "r"(row0), "r"(s_u32(bar)) : "memory");
}

// S[g][c][t] = X[t][c][g] (t < T, zero up to Ts): reads coalesced along g, writes coalesced along t
__global__ void k_cov_gather(PerBinArgs a) {
  __shared__ float2 tile[32][33];
  const int c = blockIdx.y;
  const int g0 = blockIdx.x * 32, t0 = blockIdx.z * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, g = g0 + threadIdx.x;
    tile[j][threadIdx.x] = (t < a.T && g < a.G) ? a.X[((size_t)t * a.C + c) * a.Gp + g] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int g = g0 + j, t = t0 + threadIdx.x;
    if (g < a.G && t < a.Ts) a.Scov[((size_t)g * a.C + c) * a.Ts + t] = tile[threadIdx.x][j];
  }
}

__global__ void __launch_bounds__(THREADS, 1) k_covariance_tc(const __grid_constant__ CUtensorMap tmX, PerBinArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment by OFFSET from the __shared__ symbol, so the compiler keeps the shared address space (LDS / STS): with a
  // uintptr_t round trip every access became a generic LD.E / ST.E, which the mbarrier arrive below does not wait for
  unsigned char* base = smem_raw + ((1024u - (s_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* raw = base + (size_t)NSTAGE * STAGE_B;   // [NRAW][RAW_B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + (size_t)NRAW * RAW_B);
  uint64_t* full = bars;              // [NSTAGE] transposers -> MMA
  uint64_t* empty = bars + NSTAGE;    // [NSTAGE] MMA (tcgen05.commit) -> transposers
  uint64_t* accf = bars + 2 * NSTAGE; // [2] MMA -> epilogue
  uint64_t* acce = accf + 2;          // [2] epilogue -> MMA
  uint64_t* rfull = acce + 2;         // [NRAW] TMA -> transposers
  uint64_t* rempty = rfull + NRAW;    // [NRAW] transposers -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + NRAW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; s++) { mb_init(full + s, PROD_WARPS); mb_init(empty + s, 1); }   // one arrival per transposer warp
    for (int b = 0; b < 2; b++) { mb_init(accf + b, 1); mb_init(acce + b, EPI_WARPS * 32); }
    for (int s = 0; s < NRAW; s++) { mb_init(rfull + s, 1); mb_init(rempty + s, PROD_WARPS); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int G = a.G, T = a.T, K = a.K, U = a.U;
  const int ngroups = (G + NCH - 1) / NCH;
  const int NKB = (T + KB - 1) / KB;

  // The pipeline unit is ONE chain's K-block (it = ((group, kb), j), j fastest): 4 operand stages and 6 raw stages in the same
  // 224 KiB give twice the depth of pair-sized stages, which this latency-bound pipeline needs.
  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------------------------------------ raw gather (TMA)
    if (lane == 0) {
      int it = 0;
      for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        for (int kb = 0; kb < NKB; kb++) {
          for (int j = 0; j < NCH; j++, it++) {
            const int rs = it % NRAW;
            mb_wait(rempty + rs, (uint32_t)(((it / NRAW) & 1) ^ 1));
            mb_expect_tx(rfull + rs, (uint32_t)RAW_B);
            tma_load_2d(raw + (size_t)rs * RAW_B, &tmX, 2 * kb * KB, (NCH * grp + j) * C64, rfull + rs);   // rows past G C: zero fill
          }
        }
      }
    }
    __syncwarp();
  } else if (warp > MMA_WARP) {
    // ------------------------------------------------------------------------------------------------ transposers
    const int pw = (warp - MMA_WARP - 1) & 7;    // (channel & 7) of every row this warp writes
    const int ph = (warp - MMA_WARP - 1) >> 3;   // 0 / 1: which half of the 8-channel blocks (PROD_WARPS = 16)
    constexpr int NI = 64 / PROD_WARPS;          // channels per warp
    int it = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
      const int g0 = grp * NCH;
      for (int kb = 0; kb < NKB; kb++) {
        const int t = kb * KB + lane;
        for (int j = 0; j < NCH; j++, it++) {
          const int s = it % NSTAGE, rs = it % NRAW;
          const bool vj = g0 + j < G;
          const bool m = vj && t < T && a.noise_mask[(size_t)t * U + (g0 + j) / K] != 0;
          mb_wait(rfull + rs, (uint32_t)((it / NRAW) & 1));
          float2 v[NI];
          const unsigned char* rb = raw + (size_t)rs * RAW_B;
#pragma unroll
          for (int i = 0; i < NI; i++) {
            const int cb = (PROD_WARPS == 16) ? 2 * i + ph : i;   // 8-channel block of this iteration
            const float2 x0 = *reinterpret_cast<const float2*>(rb + ((size_t)(pw + 8 * cb) * KB + lane) * 8);
            v[i] = make_float2(m ? x0.x : 0.f, m ? x0.y : 0.f);
          }
          mb_wait(empty + s, (uint32_t)(((it / NSTAGE) & 1) ^ 1));
          unsigned char* th = base + (size_t)s * STAGE_B;
          unsigned char* tl = th + TILE_B;
#pragma unroll
          for (int i = 0; i < NI; i++) {
            // channel c = pw + 8 cb: Xr row 16 cb + pw, Xi row 16 cb + 8 + pw; column = lane (frame within the K-block)
            const int cb = (PROD_WARPS == 16) ? 2 * i + ph : i;
            const uint32_t off1 = (uint32_t)(16 * cb + pw) * 128u + ((uint32_t)((lane >> 2) ^ pw) << 4) + ((uint32_t)(lane & 3) << 2);
            const uint32_t off2 = off1 + 8u * 128u;
            const float xr = v[i].x, xi = v[i].y;
            const float hr = __uint_as_float(__float_as_uint(xr) & 0xffffe000u), hi = __uint_as_float(__float_as_uint(xi) & 0xffffe000u);
            *reinterpret_cast<float*>(th + off1) = hr; *reinterpret_cast<float*>(tl + off1) = xr - hr;
            *reinterpret_cast<float*>(th + off2) = hi; *reinterpret_cast<float*>(tl + off2) = xi - hi;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's async proxy
          __syncwarp();                                                   // every lane's stores are fenced before the warp's one arrival
          if (lane == 0) mb_arrive(full + s);
          // the raw slot is released only now: its values have been consumed by the stores above, so the loads have certainly
          // completed before the TMA unit (async proxy) may overwrite the slot.  Releasing right after ISSUING the loads let the
          // refill race them (sporadic 10-20 % errors in fully masked K-blocks, where the transposers run ahead).
          if (lane == 0) mb_arrive(rempty + rs);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int it = 0, gi = 0;
      for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x, gi++) {
        const int b = gi & 1;
        mb_wait(acce + b, (uint32_t)(((gi >> 1) & 1) ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < NKB; kb++) {
          for (int j = 0; j < NCH; j++, it++) {
            const int s = it % NSTAGE;
            mb_wait(full + s, (uint32_t)((it / NSTAGE) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = s_u32(base + (size_t)s * STAGE_B), al = ah + TILE_B;
            const uint32_t d = tmem_base + (uint32_t)(b * NCH * 128 + j * 128);
#pragma unroll
            for (int ks = 0; ks < KB / 8; ks++) {
              const uint64_t dh = smem_desc(ah + ks * 32), dl = smem_desc(al + ks * 32);
              umma_tf32(d, dh, dh, (kb > 0 || ks > 0) ? 1u : 0u);   // hi hi^T
              umma_tf32(d, dl, dh, 1u);                               // lo hi^T
              umma_tf32(d, dh, dl, 1u);                               // hi lo^T
            }
            umma_commit(empty + s);   // the slot may be refilled once these MMAs have read it
          }
        }
        umma_commit(accf + b);      // accumulators of this group complete
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue
    const int r = 32 * warp + lane;                  // TMEM lane = operand row
    const int c = 8 * (r >> 4) + (r & 7);            // channel of this row
    const bool is_xi = ((r >> 3) & 1) != 0;
    int gi = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x, gi++) {
      const int b = gi & 1;
      const int g0 = grp * NCH;
      mb_wait(accf + b, (uint32_t)((gi >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int j = 0; j < NCH; j++) {
        if (g0 + j >= G) break;   // uniform over the CTA
        float2* Rg = a.R + (size_t)(g0 + j);
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
          uint32_t rr[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(b * NCH * 128 + j * 128 + q * 32), rr);
          float val[16];
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const int xr_reg = 16 * (e >> 3) + (e & 7), xi_reg = xr_reg + 8;
            const float own = __uint_as_float(rr[xr_reg]);
            const float got = __shfl_xor_sync(0xffffffffu, __uint_as_float(rr[xi_reg]), 8);
            val[e] = is_xi ? own - got : own + got;   // Xr lane: Re R[c][c'], Xi lane: Im R[c][c'], c' = 16 q + e
          }
#pragma unroll
          for (int sft = 0; sft < 8; sft++) {
            const float send = is_xi ? val[sft] : val[8 + sft];
            const float got = __shfl_xor_sync(0xffffffffu, send, 8);
            const int cp = 16 * q + (is_xi ? 8 + sft : sft);
            const float2 out = is_xi ? make_float2(got, val[8 + sft]) : make_float2(val[sft], got);
            Rg[(size_t)(c * C64 + cp) * a.Gp] = out;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mb_arrive(acce + b);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc

// S viewed as float32 [G C rows][2 Ts]; box {2 KB floats, 64 rows}
static cudaError_t make_series_map(CUtensorMap* tm, const PerBinArgs& a) {
  return encode_tensor_map_2d_f32(tm, a.Scov, (cuuint64_t)2 * a.Ts, (cuuint64_t)a.G * a.C, (cuuint64_t)a.Ts * sizeof(float2), (cuuint32_t)(2 * tc::KB),
                                  (cuuint32_t)tc::C64, CU_TENSOR_MAP_SWIZZLE_NONE);
}

size_t covariance_tc_workspace_bytes(int G, int C, int T) { return (size_t)G * C * (size_t)((T + tc::KB - 1) / tc::KB * tc::KB) * sizeof(float2); }

// a.Scov: workspace of covariance_tc_workspace_bytes(G, C, T) bytes
cudaError_t launch_covariance_tc(const PerBinArgs& a_in, cudaStream_t st) {
  if (a_in.T <= 0 || a_in.G <= 0) return cudaSuccess;
  if (a_in.C != tc::C64 || a_in.Scov == nullptr) return cudaErrorInvalidValue;
  PerBinArgs a = a_in;
  a.Ts = (a.T + tc::KB - 1) / tc::KB * tc::KB;
  {
    dim3 grid((a.G + 31) / 32, a.C, a.Ts / 32), block(32, 8);
    tc::k_cov_gather<<<grid, block, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaFuncSetAttribute(tc::k_covariance_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  CUtensorMap tm;
  e = make_series_map(&tm, a);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ngroups = (a.G + tc::NCH - 1) / tc::NCH;
  const int grid = ngroups < sms ? ngroups : sms;   // persistent: one CTA per SM (224 KiB of shared memory, all 512 TMEM columns)
  tc::k_covariance_tc<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(tm, a);
  return cudaGetLastError();
}

}  // namespace btkb
