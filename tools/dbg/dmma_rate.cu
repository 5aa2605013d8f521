// Micro-benchmark: issue rate of the fp64 tensor-core MMA (mma.sync m8n8k4 f64) against plain DFMA on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(a, b, c[i]);
  }
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = warps * 32, blocks = 148;
    float ms;
    k_dmma<8><<<blocks, threads>>>(out, 100, 1.0, 1.0); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dmma<8><<<blocks, threads>>>(out, iters, 1.0, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    const double fl_mma = 2.0 * 256 * 8 * (double)iters * warps * blocks;
    printf("DMMA m8n8k4: %2d warps/SM  %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl_mma / ms * 1e-9);
    k_dfma<16><<<blocks, threads>>>(out, 100, 1.0, 1.0); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dfma<16><<<blocks, threads>>>(out, iters, 1.0, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    const double fl_fma = 2.0 * 32 * 16 * (double)iters * warps * blocks;
    printf("DFMA       : %2d warps/SM  %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl_fma / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
