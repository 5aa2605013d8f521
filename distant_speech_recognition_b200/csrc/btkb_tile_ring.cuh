// btkb_tile_ring.cuh — the tensor-map TMA shared-memory ring every per-chain kernel (btkb_perbin.cu, btkb_sos.cu) reads X through.
#pragma once
#include <cuda.h>           // CUtensorMap (the encode entry point is fetched at run time; libcuda is not linked)
#include "btkb_internal.h"
#include "btkb_tensor_map.h"

namespace btkb {

constexpr int TILE = 64;    // chains (threads) per CTA: 1028 CTAs at configs[1] = 6.95 per SM (balanced single wave)
constexpr int FCH = 2;      // frames per ring slot: one 2-D TMA box is [FCH*C rows][TILE chains]
constexpr int STAGES = 3;   // ring slots per CTA (2 slots = 4 frames in flight); 24 KiB per CTA so 7 CTAs fit one SM (single balanced wave)
constexpr int NWARP = TILE / 32;

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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 2-D tiled TMA load: box {2*TILE floats, FCH*C rows} of the [T*C][2*Gp] float view of X at (col0, row0) -> smem
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int col0, int row0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(col0), "r"(row0), "r"(smem_u32(bar)) : "memory");
}

// Ring of STAGES slots; a slot holds FCH consecutive frames of the CTA's mic x bin tile, [FCH][C][TILE] complex64, and is
// filled by ONE tensor-map TMA instruction (rows (t, c) of X are consecutive, so FCH frames x C channels is a dense 2-D
// box; rows past the end of the batch are zero-filled by the TMA unit).  full[s]: armed with the slot's byte count,
// completed by the TMA.  empty[s]: one arrival per warp once all its lanes hold the slot's last frame in registers.
// The producer (lane 0 of warp 0) refills the previous slot while the CTA works on the current one, so warps never meet
// at a CTA-wide barrier.
template <int C>
struct TileRing {
  float2* stage;     // [STAGES][FCH][C][TILE]
  uint64_t* full;    // [STAGES]
  uint64_t* empty;   // [STAGES]
  const CUtensorMap* tm; int g0, NS;   // NS = number of slots' worth of frames = ceil(T / FCH)
  static constexpr uint32_t SLOT_BYTES = FCH * C * TILE * sizeof(float2);
  __device__ __forceinline__ void init(unsigned char* smem, const CUtensorMap* tm_, int g0_, int T_) {
    stage = reinterpret_cast<float2*>(smem);
    full = reinterpret_cast<uint64_t*>(smem + (size_t)SLOT_BYTES * STAGES);
    empty = full + STAGES;
    tm = tm_; g0 = g0_; NS = (T_ + FCH - 1) / FCH;
    if (threadIdx.x == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
      for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, NWARP); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int j = 0; j < STAGES - 1 && j < NS; j++) issue(j);
  }
  // called by ONE thread: fetch frame group j (frames j*FCH ..) into slot j % STAGES
  __device__ __forceinline__ void issue(int j) {
    const int s = j % STAGES;
    mbar_expect_tx(full + s, SLOT_BYTES);
    tma_load_2d(reinterpret_cast<unsigned char*>(stage) + (size_t)s * SLOT_BYTES, tm, 2 * g0, j * FCH * C, full + s);
  }
  // all threads, at the first frame of group j: the producer refills the previous slot, everybody waits for this one
  __device__ __forceinline__ void acquire(int j) {
    if (threadIdx.x == 0 && j + STAGES - 1 < NS) {
      if (j >= 1) mbar_wait(empty + ((j - 1) % STAGES), (uint32_t)(((j - 1) / STAGES) & 1));
      issue(j + STAGES - 1);
    }
    mbar_wait(full + (j % STAGES), (uint32_t)((j / STAGES) & 1));
  }
  __device__ __forceinline__ void load(int j, int f, float2* x) const {
    const float2* base = stage + ((size_t)(j % STAGES) * FCH + f) * C * TILE + threadIdx.x;
#pragma unroll
    for (int c = 0; c < C; c++) x[c] = base[(size_t)c * TILE];
  }
  // all threads, after loading the last frame of group j into registers
  __device__ __forceinline__ void release(int j) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty + (j % STAGES));
  }
  // frame t -> registers (wraps acquire / load / release)
  __device__ __forceinline__ void fetch(int t, int T, float2* x) {
    const int j = t / FCH, f = t % FCH;
    if (f == 0) acquire(j);
    load(j, f, x);
    if (f == FCH - 1 || t == T - 1) release(j);
  }
};

// bytes of dynamic shared memory the ring owns (slots + mbarriers, padded to 128 so what follows stays aligned)
template <int C>
__host__ __device__ constexpr size_t ring_bytes() { return ((sizeof(float2) * STAGES * FCH * C * TILE + sizeof(uint64_t) * 2 * STAGES + 64) + 127) / 128 * 128; }


// Tensor map of X viewed as a 2-D float tensor [T*C rows][2*Gp floats]; box = [FCH*C rows][2*TILE floats].
static inline cudaError_t make_tensor_map(CUtensorMap* tm, const PerBinArgs& a, int C) {
  return encode_tensor_map_2d_f32(tm, a.X, (cuuint64_t)2 * a.Gp, (cuuint64_t)a.T * C, (cuuint64_t)a.Gp * sizeof(float2), (cuuint32_t)(2 * TILE),
                                  (cuuint32_t)(FCH * C), CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace btkb
