// btkb_tensor_map.h — one place for the tensor-map (TMA descriptor) encoding every TMA-fed kernel uses.  libcuda is not linked:
// cuTensorMapEncodeTiled is fetched through the runtime at first use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace btkb {

// 2-D float32 view: gdim0 x gdim1 elements, row pitch `row_bytes`, box box0 x box1, unit element strides, no OOB fill value (zeros)
static inline cudaError_t encode_tensor_map_2d_f32(CUtensorMap* tm, const void* base, cuuint64_t gdim0, cuuint64_t gdim1, cuuint64_t row_bytes,
                                                   cuuint32_t box0, cuuint32_t box1, CUtensorMapSwizzle swizzle,
                                                   CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return cudaErrorNotSupported;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  cuuint64_t gdim[2] = {gdim0, gdim1};
  cuuint64_t gstride[1] = {row_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (r == CUDA_SUCCESS) ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace btkb
