// Latency / issue-interval micro-benchmarks of the instructions the small dense solver is made of (one warp, one SM):
// DFMA, MUFU.RCP64H, MUFU.RSQ64H + Newton, SHFL of a double, DMMA m8n8k4 dependent chain and independent streams.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat tools/micro/lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__global__ void k(double* out, long long* cyc, double seed) {
  const int lane = threadIdx.x & 31;
  double x = seed + lane * 1e-3, y = 1.0 + 1e-9 * lane;
  constexpr int N = 256;
  long long t0, t1;
  // 0: dependent DFMA
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
  t1 = clk(); if (threadIdx.x == 0) cyc[0] = (t1 - t0); out[threadIdx.x] = x;
  // 1: dependent rcp.approx.ftz.f64
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  t1 = clk(); if (threadIdx.x == 0) cyc[1] = (t1 - t0); out[threadIdx.x] += x;
  x = seed + lane * 1e-3;
  // 2: dependent rsqrt()
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = rsqrt(x) + 1.0;
  t1 = clk(); if (threadIdx.x == 0) cyc[2] = (t1 - t0); out[threadIdx.x] += x;
  // 3: dependent double shuffle
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (lane + 5) & 31);
  t1 = clk(); if (threadIdx.x == 0) cyc[3] = (t1 - t0); out[threadIdx.x] += x;
  // 4: dependent DMMA chain
  double c0 = 0, c1 = 0;
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) dmma(c0, c1, y, x);
  t1 = clk(); if (threadIdx.x == 0) cyc[4] = (t1 - t0); out[threadIdx.x] += c0 + c1;
  // 5: 8 independent DMMA streams (per warp)
  double a0[8], a1[8];
#pragma unroll
  for (int s = 0; s < 8; s++) { a0[s] = s; a1[s] = -s; }
  t0 = clk();
#pragma unroll 4
  for (int i = 0; i < N / 8; i++) {
#pragma unroll
    for (int s = 0; s < 8; s++) dmma(a0[s], a1[s], y, x);
  }
  t1 = clk(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
#pragma unroll
  for (int s = 0; s < 8; s++) out[threadIdx.x] += a0[s] + a1[s];
  // 6: DMMA result -> dependent through the A operand (TRSM pattern: x = A M; next uses x as operand)
  c0 = 1; c1 = 2;
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) { double d0 = 0, d1 = 0; dmma(d0, d1, c0, x); c0 = d0; }
  t1 = clk(); if (threadIdx.x == 0) cyc[6] = (t1 - t0); out[threadIdx.x] += c0;
  // 7: dependent 1.0 / x (IEEE) and __drcp_rn
  x = seed + lane * 1e-3;
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __drcp_rn(x) + 0.5;
  t1 = clk(); if (threadIdx.x == 0) cyc[7] = (t1 - t0); out[threadIdx.x] += x;
  // 8: dependent LDS of a double
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = (double)((threadIdx.x * 7) & 31);
  __syncthreads();
  int idx = lane;
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) idx = (int)sm[idx];
  t1 = clk(); if (threadIdx.x == 0) cyc[8] = (t1 - t0); out[threadIdx.x] += idx;
  // 9: __syncthreads round trip with blockDim threads
  t0 = clk();
#pragma unroll 16
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clk(); if (threadIdx.x == 0) cyc[9] = (t1 - t0);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
  const char* names[10] = {"DFMA dependent", "rcp.approx.ftz.f64 dependent", "rsqrt(double)+add dependent", "shfl double dependent",
                           "DMMA m8n8k4 dependent (accumulator)", "DMMA 8 independent streams (per DMMA)", "DMMA dependent through A operand",
                           "__drcp_rn + add dependent", "LDS double -> index dependent", "__syncthreads"};
  for (int threads : {32, 512}) {
    for (int rep = 0; rep < 2; rep++) k<<<1, threads>>>(out, cyc, 1.2345);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    printf("threads per CTA %d (cycles per operation, warp 0):\n", threads);
    for (int i = 0; i < 10; i++) printf("  %-44s %7.1f\n", names[i], h[i] / 256.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
