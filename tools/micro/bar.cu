#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ double s[4160];
  __shared__ double inv[80];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4160; i += blockDim.x) s[i] = 1.0 + 1e-3 * i;
  if (tid < 80) inv[tid] = 1.0 + tid;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int j = it & 63;
    const double v = inv[j];
    if (MODE >= 1) {                 // thread 0: dependent rcp + store
      if (tid == 0) inv[j + 1] = __drcp_rn(v + 1.5);
    }
    if (MODE >= 2) {                 // every thread: 2 read-modify-writes depending on v
      const int a = (tid * 2) & 4095;
      s[a] = s[a] - v * s[(a + 65) & 4095];
      s[a + 1] = s[a + 1] - v * s[(a + 66) & 4095];
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[tid] = s[tid] + inv[tid & 63];
  if (tid == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 64);
  const int iters = 6400;
  for (int threads : {256, 512, 1024}) {
    k<0><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("%4d threads: LDS + barrier                 %.1f cycles/iter\n", threads, (double)cyc[0] / iters);
    k<1><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("%4d threads: + thread-0 rcp chain            %.1f cycles/iter\n", threads, (double)cyc[0] / iters);
    k<2><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("%4d threads: + 2 smem RMW per thread         %.1f cycles/iter\n", threads, (double)cyc[0] / iters);
  }
  return 0;
}
