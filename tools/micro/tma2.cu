// Controls for tma.cu: 6 = libcu++ cde:: 2D tensor copy (the programming guide's example), 7 = 1D bulk copy without descriptor
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void k6(const __grid_constant__ CUtensorMap tensor_map, uint8_t* out, int x, int y) {
  __shared__ alignas(128) uint8_t in[38 * 144];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&in, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(in));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 38 * 144; i += blockDim.x) out[i] = in[i];
}
__global__ void k7(const uint8_t* src, uint8_t* out) {
  __shared__ alignas(128) uint8_t in[4096];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cuda::memcpy_async(in, src, cuda::aligned_size_t<16>(4096), bar);
    token = bar.arrive();
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = in[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 6;
  const int pitch = 1408, rows = 414;
  uint8_t *d, *out; cudaMalloc(&d, (size_t)pitch * rows); cudaMalloc(&out, 38 * 144);
  uint8_t* hbuf = new uint8_t[(size_t)pitch * rows];
  for (size_t i = 0; i < (size_t)pitch * rows; i++) hbuf[i] = (uint8_t)(i * 2654435761u >> 24);
  cudaMemcpy(d, hbuf, (size_t)pitch * rows, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap one;
  const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {144, 38}, estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&one, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (variant == 6) k6<<<1, 256>>>(one, out, 28, 16); else k7<<<1, 256>>>(d, out);
  cudaError_t s = cudaDeviceSynchronize();
  printf("variant %d: encode %d, run: %s", variant, (int)r, cudaGetErrorString(s));
  if (s == cudaSuccess) {
    uint8_t hout[38 * 144];
    cudaMemcpy(hout, out, variant == 6 ? 38 * 144 : 4096, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (variant == 6) { for (int rr = 0; rr < 38; rr++) for (int c = 0; c < 144; c++) bad += hout[rr * 144 + c] != hbuf[(size_t)(16 + rr) * pitch + 28 + c]; }
    else for (int i = 0; i < 4096; i++) bad += hout[i] != hbuf[i];
    printf(", %d mismatches", bad);
  }
  printf("\n");
  return 0;
}
