#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void kdep(double* out, long long* cyc, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.0;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c); }
  long long t1 = clock64();
  out[threadIdx.x] = a0;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void kdiv(double* out, long long* cyc, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a0 = 1.0 / a0 + 0.5; }
  long long t1 = clock64();
  double a1 = threadIdx.x * 1e-9 + 1.5;
  for (int i = 0; i < iters; i++) { a1 = __drcp_rn(a1) + 0.5; }
  long long t2 = clock64();
  double a2 = threadIdx.x * 1e-9 + 1.5;
  for (int i = 0; i < iters; i++) { a2 = rsqrt(a2) + 0.5; }
  long long t3 = clock64();
  out[threadIdx.x] = a0 + a1 + a2;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 4096 * 8);
  const int iters = 4096;
  for (int threads : {32, 128, 256, 512, 1024}) {
    k<<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("1 CTA x %4d threads: %lld cycles, %.2f cycles per warp-DFMA per SM (8 indep chains)\n", threads, cyc[0], (double)cyc[0] / (iters * 8.0 * (threads / 32)));
  }
  k<<<148, 1024>>>(out, cyc, iters); cudaDeviceSynchronize();
  printf("148 CTAs x 1024: %lld cycles -> %.2f cycles per warp-DFMA per SM\n", cyc[0], (double)cyc[0] / (iters * 8.0 * 32));
  kdep<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  printf("dependent DFMA latency: %.1f cycles\n", (double)cyc[0] / (iters * 4.0));
  kdiv<<<1, 32>>>(out, cyc, 1024); cudaDeviceSynchronize();
  printf("dependent 1/x+c: %.1f cycles; __drcp_rn+c: %.1f; rsqrt+c: %.1f\n", cyc[0] / 1024.0, cyc[1] / 1024.0, cyc[2] / 1024.0);
  return 0;
}
