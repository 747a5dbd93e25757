// TMA smoke test (one variant per process: an illegal instruction poisons the context): ./tma <variant>
//   0 param descriptor, thread 0 issues     1 param array + dynamic index     2 elect.sync in a converged warp
//   3 descriptor in global memory            4 2D tensor                        5 box 128 x 38 (power of two inner size)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
struct Maps { CUtensorMap m[8]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}
__global__ void k(const __grid_constant__ CUtensorMap one, const __grid_constant__ Maps maps, const CUtensorMap* gmap, int variant,
                  int level, uint8_t* out, int x, int y, int z, int bytes) {
  __shared__ __align__(1024) uint8_t in[38 * 144];
  __shared__ __align__(8) uint64_t bar;
  const CUtensorMap* map = variant == 1 ? &maps.m[level] : variant == 3 ? gmap : &one;
  bool issuer = threadIdx.x == 0;
  if (variant == 2) issuer = (threadIdx.x < 32) && elect_one();
  if (issuer) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (variant == 4)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(in)), "l"(map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(in)), "l"(map), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  __syncthreads();
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = in[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 1408, rows = 414, frames = 4; const size_t frame_bytes = (size_t)pitch * rows + 256 * 3;
  uint8_t *d, *out; cudaMalloc(&d, frame_bytes * frames); cudaMalloc(&out, 38 * 144);
  uint8_t* hbuf = new uint8_t[frame_bytes * frames];
  for (size_t i = 0; i < frame_bytes * frames; i++) hbuf[i] = (uint8_t)(i * 2654435761u >> 24);
  cudaMemcpy(d, hbuf, frame_bytes * frames, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Maps maps; CUtensorMap one;
  const int bw = variant == 5 ? 128 : 144, rank = variant == 4 ? 2 : 3;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_bytes};
  const cuuint32_t box[3] = {(cuuint32_t)bw, 38, 1}, estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int l = 0; l < 8; l++)
    r = ((EncodeFn)fn)(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  one = maps.m[0];
  CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &one, sizeof(one), cudaMemcpyHostToDevice);
  const int x = 28, y = 16, z = rank == 2 ? 0 : 2, bytes = 38 * bw;
  k<<<1, 256>>>(one, maps, gmap, variant, 3, out, x, y, z, bytes);
  cudaError_t s = cudaDeviceSynchronize();
  printf("variant %d: encode %d, run: %s", variant, (int)r, cudaGetErrorString(s));
  if (s == cudaSuccess) {
    uint8_t hout[38 * 144];
    cudaMemcpy(hout, out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < 38; rr++) for (int c = 0; c < bw; c++) bad += hout[rr * bw + c] != hbuf[z * frame_bytes + (size_t)(y + rr) * pitch + x + c];
    printf(", %d mismatches", bad);
  }
  printf("\n");
  return 0;
}
