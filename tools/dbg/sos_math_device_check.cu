// Debug aid: run the stages of sos_solve_chain (csrc/btkb_sos_math.cuh) on the host and on the device for the same random problems.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../distant_speech_recognition_b200/csrc/btkb_sos_math.cuh"
using namespace btkb;

// out layout per problem: Cm[C*C], L[C*C], y[C], w[C], lambda (as zd)
template <int C>
__host__ __device__ void stages(const zd* RtA, const zd* RnA, zd* out, int kind) {
  zd Rt[C][C], Rn[C][C], w[C], y[C];
  for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) { Rt[i][j] = RtA[i * C + j]; Rn[i][j] = RnA[i * C + j]; }
  for (int c = 0; c < C; c++) { w[c] = zmk(0, 0); y[c] = zmk(0, 0); }
  double lam = 0;
  if (kind == 0) { improve_condition<C>(Rn, 1e-6); bmvdr_solve<C>(Rt, Rn, 0, 0.0, w); }
  else {
    improve_condition<C>(Rn, 1e-6);
    gev_reduce<C>(Rt, Rn);
    for (int i = 0; i < C * C; i++) { out[i] = Rt[i / C][i % C]; out[C * C + i] = Rn[i / C][i % C]; }
    lam = jacobi_principal<C>(Rt, y);
    gev_back<C>(Rn, y, w);
  }
  for (int c = 0; c < C; c++) { out[2 * C * C + c] = y[c]; out[2 * C * C + C + c] = w[c]; }
  out[2 * C * C + 2 * C] = zmk(lam, 0);
}
template <int C>
__global__ void k(const zd* RtA, const zd* RnA, zd* out, int n, int kind) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) stages<C>(RtA + (size_t)q * C * C, RnA + (size_t)q * C * C, out + (size_t)q * (2 * C * C + 2 * C + 1), kind);
}

static double rnd() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static double rel(const zd* a, const zd* b, int n, bool lower_only = false, int C = 0) {
  double num = 0, den = 0;
  for (int i = 0; i < n; i++) { if (lower_only && (i % C) > (i / C)) continue; zd d = zsub(a[i], b[i]); num += zabs2(d); den += zabs2(b[i]); }
  return sqrt(num / (den > 0 ? den : 1));
}

template <int C>
void run(int kind, double scale) {
  const int n = 64, S = 2 * C * C + 2 * C + 1;
  std::vector<zd> Rt((size_t)n * C * C), Rn((size_t)n * C * C), Oh((size_t)n * S), Od((size_t)n * S);
  for (int q = 0; q < n; q++) {
    auto herm = [&](zd* A, int rank, double sc, double load) {
      for (int i = 0; i < C * C; i++) A[i] = zmk(0, 0);
      for (int r = 0; r < rank; r++) {
        zd v[C];
        for (int c = 0; c < C; c++) v[c] = zmk(rnd(), rnd());
        for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) A[i * C + j] = zadd(A[i * C + j], zscale(zmulc(v[i], v[j]), sc));
      }
      for (int i = 0; i < C; i++) A[i * C + i].x += load;
    };
    herm(&Rt[(size_t)q * C * C], 3, scale, 0.0);
    herm(&Rn[(size_t)q * C * C], 4 * C, 1.0, 0.1);
  }
  for (int q = 0; q < n; q++) stages<C>(&Rt[(size_t)q * C * C], &Rn[(size_t)q * C * C], &Oh[(size_t)q * S], kind);
  zd *dRt, *dRn, *dO;
  cudaMalloc(&dRt, Rt.size() * sizeof(zd)); cudaMalloc(&dRn, Rn.size() * sizeof(zd)); cudaMalloc(&dO, Od.size() * sizeof(zd));
  cudaMemset(dO, 0, Od.size() * sizeof(zd));
  cudaMemcpy(dRt, Rt.data(), Rt.size() * sizeof(zd), cudaMemcpyHostToDevice); cudaMemcpy(dRn, Rn.data(), Rn.size() * sizeof(zd), cudaMemcpyHostToDevice);
  k<C><<<1, 64>>>(dRt, dRn, dO, n, kind);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(Od.data(), dO, Od.size() * sizeof(zd), cudaMemcpyDeviceToHost);
  double wCm = 0, wL = 0, wy = 0, ww = 0, wl = 0;
  for (int q = 0; q < n; q++) {
    const zd* h = &Oh[(size_t)q * S]; const zd* d = &Od[(size_t)q * S];
    wCm = fmax(wCm, rel(d, h, C * C)); wL = fmax(wL, rel(d + C * C, h + C * C, C * C, true, C));
    wy = fmax(wy, rel(d + 2 * C * C, h + 2 * C * C, C)); ww = fmax(ww, rel(d + 2 * C * C + C, h + 2 * C * C + C, C));
    wl = fmax(wl, fabs(d[2 * C * C + 2 * C].x - h[2 * C * C + 2 * C].x) / fabs(h[2 * C * C + 2 * C].x + 1e-300));
  }
  printf("C=%d kind=%d scale=%g: %s  worst device-vs-host: Cm %.2e  L %.2e  lambda %.2e  y %.2e  w %.2e\n", C, kind, scale, cudaGetErrorString(e), wCm, wL, wl, wy, ww);
  cudaFree(dRt); cudaFree(dRn); cudaFree(dO);
}

int main() {
  srand(1);
  run<4>(0, 50.0); run<2>(1, 50.0); run<4>(1, 50.0); run<8>(1, 50.0); run<8>(1, 1e10);
  return 0;
}
