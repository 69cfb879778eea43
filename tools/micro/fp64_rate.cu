// dev microbenchmark: sustained DFMA rate per SM with N independent accumulators and W warps per SMSP
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k_dfma(double* out, int iters, double x, double y) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DFMA fed from shared memory like K1a: MR + 4 LDS.64 per MR*4 DFMA
template <int MR>
__global__ void k_tile(double* out, int iters, int ld) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 104 * ld; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int tx = threadIdx.x % 13, ty = (threadIdx.x / 13) % 8;
  double acc[MR][4];
#pragma unroll
  for (int r = 0; r < MR; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0;
  const double* pa = sm + ty * ld;
  const double* pb = sm + (MR * 8 + tx) * ld;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      double a[MR], b[4];
#pragma unroll
      for (int r = 0; r < MR; ++r) a[r] = pa[r * 8 * ld + k];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = pb[q * 13 * ld + k];
#pragma unroll
      for (int r = 0; r < MR; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = fma(a[r], b[q], acc[r][q]);
    }
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < MR; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) s += acc[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma(double* out, int iters, double x, double y) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(x), "d"(y));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* d; cudaMalloc(&d, 148 * 16 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps_per_sm : {4, 8, 16, 32}) {
    const int threads = 128, blocks = 148 * warps_per_sm / 4, iters = 20000;
    k_dfma<20><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k_dfma<20><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * blocks * threads * (double)iters * 20;
    printf("pure DFMA, %2d warps/SM: %.1f TFLOP/s\n", warps_per_sm, fl / ms / 1e9);
  }
  for (int warps_per_sm : {4, 8, 16}) {
    const int threads = 128, blocks = 148 * warps_per_sm / 4, iters = 20000;
    k_dmma<7><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k_dmma<7><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * blocks * (threads / 32) * (double)iters * 7 * 256;
    printf("pure DMMA m8n8k4, %2d warps/SM: %.1f TFLOP/s\n", warps_per_sm, fl / ms / 1e9);
  }
  for (int ld : {17}) for (int ctas : {4}) {
    const int iters = 3000, blocks = 148 * ctas;
    cudaFuncSetAttribute(k_tile<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
    k_tile<5><<<blocks, 128, 104 * ld * 8>>>(d, 10, ld);
    cudaEventRecord(e0);
    k_tile<5><<<blocks, 128, 104 * ld * 8>>>(d, iters, ld);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * blocks * 128 * (double)iters * 16 * 20;
    printf("smem-fed 5x4 tile, ld %d, %d CTAs/SM (128 thr): %.1f TFLOP/s (%.3f ms)\n", ld, ctas, fl / ms / 1e9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
