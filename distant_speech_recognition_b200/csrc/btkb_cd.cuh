// btkb_cd.cuh — fp64 complex helpers and the blocking matrix shared by the setup kernels (btkb_weights.cu) and the fp64 recurrence of
// the reference's C++ SubbandGSCRLS (btkb_rls_cpp.cu).
#pragma once
#include <cuda_runtime.h>

namespace btkb {

struct cd { double x, y; };
__device__ __forceinline__ cd cdmake(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cd cdadd(cd a, cd b) { return cdmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd cdsub(cd a, cd b) { return cdmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cdmul(cd a, cd b) { return cdmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cdconj(cd a) { return cdmake(a.x, -a.y); }
__device__ __forceinline__ cd cdscale(cd a, double s) { return cdmake(a.x * s, a.y * s); }
__device__ __forceinline__ double cdabs2(cd a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ cd cddiv(cd a, cd b) { double d = cdabs2(b); return cdmake((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d); }

// B[C][C-1]: P = I - conj(v) v^T / ||v||^2 ; Gram-Schmidt over its first C-1 columns (NC = 1)
template <int C>
__device__ void blocking_matrix(const cd* v, cd (*B)[C - 1], int NC = 1) {
  double nv = 0.0;
  for (int c = 0; c < C; c++) nv += cdabs2(v[c]);
  cd vec[C];
  for (int idim = 0; idim < C - NC; idim++) {
    for (int r = 0; r < C; r++) {
      cd p = cdscale(cdmul(cdconj(v[r]), v[idim]), -1.0 / nv);  // P[r][idim]
      if (r == idim) p.x += 1.0;
      vec[r] = p;
    }
    for (int jdim = 0; jdim < idim; jdim++) {
      cd ip = cdmake(0, 0);
      for (int r = 0; r < C; r++) ip = cdadd(ip, cdmul(cdconj(B[r][jdim]), vec[r]));  // zdotc(rvec, vec)
      for (int r = 0; r < C; r++) vec[r] = cdsub(vec[r], cdmul(ip, B[r][jdim]));
    }
    double nrm = 0.0;
    for (int r = 0; r < C; r++) nrm += cdabs2(vec[r]);
    nrm = 1.0 / sqrt(nrm);
    for (int r = 0; r < C; r++) B[r][idim] = cdscale(vec[r], nrm);
  }
}

}  // namespace btkb
