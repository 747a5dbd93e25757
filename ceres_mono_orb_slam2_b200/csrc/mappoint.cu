// MapPoint maintenance for sm_100a, batched over map points: ComputeDistinctiveDescriptors and
// UpdateNormalAndDepth, which the reference calls for every touched point after tracking, fusing and bundle
// adjustment (src/MapPoint.cc:256-315,335-378; callers in LocalMapping.cc, LoopClosing.cc, CeresOptimizer.cc:585-598).
//
//  * k_distinctive: one warp per map point.  Row i of the all-pairs Hamming matrix is produced lane-parallel into
//    shared memory and its median (sorted row at index (int)(0.5 * (N - 1)), own zero included) is found by a
//    9-step bisection on the value range 0..256 with warp-wide counts — no sort.  First minimum median wins.
//  * k_normal_depth: one thread per map point; the normal is summed in the observation order given (the reference sums
//    in std::map order; fp64 addition is order dependent), min/max distance from the reference keyframe.
// Compiled with --fmad=false like the other front-end units.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "cmos_common.h"

namespace cmos {

constexpr int kMapMaxObs = 1024;     // observations per point handled by one warp's shared row
constexpr int kMapWarps = 4;

__global__ void __launch_bounds__(kMapWarps * 32) k_distinctive(int n_points, const int* __restrict__ obs_start,
                                                               const uint8_t* __restrict__ desc, int* __restrict__ best,
                                                               uint8_t* __restrict__ out_desc, int* __restrict__ overflow) {
  __shared__ uint16_t s_row[kMapWarps][kMapMaxObs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * kMapWarps + warp;
  if (p >= n_points) return;
  const int o = obs_start[p], N = obs_start[p + 1] - o;
  if (N <= 0) { if (lane == 0) best[p] = -1; return; }
  if (N > kMapMaxObs) { if (lane == 0) { best[p] = -1; atomicExch(overflow, 1); } return; }
  uint16_t* row = s_row[warp];
  const int k = (int)(0.5 * (N - 1));
  int best_median = 0x7fffffff, best_index = 0;
  for (int i = 0; i < N; i++) {
    uint32_t di[8];
#pragma unroll
    for (int w = 0; w < 8; w++) di[w] = __ldg((const uint32_t*)(desc + 32 * (size_t)(o + i)) + w);
    for (int j = lane; j < N; j += 32) {
      int d = 0;
      if (j != i) {
        const uint32_t* dj = (const uint32_t*)(desc + 32 * (size_t)(o + j));
#pragma unroll
        for (int w = 0; w < 8; w++) d += __popc(di[w] ^ __ldg(dj + w));
      }
      row[j] = (uint16_t)d;
    }
    __syncwarp();
    // smallest v with #{j : row[j] <= v} >= k + 1
    int lo = 0, hi = 256;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int c = 0;
      for (int j = lane; j < N; j += 32) c += row[j] <= mid;
#pragma unroll
      for (int s = 16; s; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
      if (c >= k + 1) hi = mid; else lo = mid + 1;
    }
    if (lo < best_median) { best_median = lo; best_index = i; }
    __syncwarp();
  }
  if (lane == 0) best[p] = best_index;
  if (out_desc && lane < 8)
    ((uint32_t*)(out_desc + 32 * (size_t)p))[lane] = __ldg((const uint32_t*)(desc + 32 * (size_t)(o + best_index)) + lane);
}

__global__ void __launch_bounds__(128) k_normal_depth(int n_points, const int* __restrict__ obs_start,
                                                      const int* __restrict__ obs_kf, const double* __restrict__ Ow,
                                                      const double* __restrict__ pos, const int* __restrict__ ref_kf,
                                                      const int* __restrict__ ref_level, const float* __restrict__ sf,
                                                      int n_levels, double* __restrict__ normal,
                                                      float* __restrict__ min_d, float* __restrict__ max_d) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_points) return;
  const int o = obs_start[p], N = obs_start[p + 1] - o;
  if (N <= 0) return;
  const double X[3] = {pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]};
  double nrm[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < N; k++) {
    const double* C = Ow + 3 * (size_t)obs_kf[o + k];
    const double d0 = X[0] - C[0], d1 = X[1] - C[1], d2 = X[2] - C[2];
    const double len = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    nrm[0] = nrm[0] + d0 / len; nrm[1] = nrm[1] + d1 / len; nrm[2] = nrm[2] + d2 / len;
  }
  const double* C = Ow + 3 * (size_t)ref_kf[p];
  const double c0 = X[0] - C[0], c1 = X[1] - C[1], c2 = X[2] - C[2];
  const float dist = (float)sqrt((c0 * c0 + c1 * c1) + c2 * c2);
  const float mx = dist * sf[ref_level[p]];
  max_d[p] = mx;
  min_d[p] = mx / sf[n_levels - 1];
  normal[3 * p] = nrm[0] / N; normal[3 * p + 1] = nrm[1] / N; normal[3 * p + 2] = nrm[2] / N;
}

}  // namespace cmos

using namespace cmos;

struct cmos_map {
  cmos_map_params p{};
  cudaStream_t stream = nullptr;
  int launches = 0;
  int *d_start = nullptr, *d_kf = nullptr, *d_best = nullptr, *d_ref_kf = nullptr, *d_ref_level = nullptr, *d_overflow = nullptr;
  uint8_t *d_desc = nullptr, *d_out_desc = nullptr;
  double *d_Ow = nullptr, *d_pos = nullptr, *d_normal = nullptr;
  float *d_sf = nullptr, *d_min = nullptr, *d_max = nullptr;
};

extern "C" {

int cmos_map_destroy(cmos_map_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->p.device);
  for (void* b : {(void*)h->d_start, (void*)h->d_kf, (void*)h->d_best, (void*)h->d_ref_kf, (void*)h->d_ref_level,
                  (void*)h->d_overflow, (void*)h->d_desc, (void*)h->d_out_desc, (void*)h->d_Ow, (void*)h->d_pos,
                  (void*)h->d_normal, (void*)h->d_sf, (void*)h->d_min, (void*)h->d_max})
    if (b) cudaFree(b);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CMOS_OK;
}

int cmos_map_create(const cmos_map_params* params, cmos_map_t* out) {
  CMOS_REQUIRE(params && out, "null argument");
  CMOS_REQUIRE(params->max_points > 0 && params->max_observations > 0 && params->max_keyframes > 0, "bad sizes");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_REQUIRE(params->device >= 0 && params->device < ndev, "device %d out of range", params->device);
  CMOS_CUDA_OK(cudaSetDevice(params->device));
  cmos_map* h = new cmos_map();
  h->p = *params;
  const size_t P = params->max_points, O = params->max_observations, K = params->max_keyframes;
  cudaError_t err = cudaSuccess;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) err = cudaErrorUnknown;
  h->d_start = dev_alloc<int>(P + 1, &err); h->d_kf = dev_alloc<int>(O, &err); h->d_best = dev_alloc<int>(P, &err);
  h->d_ref_kf = dev_alloc<int>(P, &err); h->d_ref_level = dev_alloc<int>(P, &err); h->d_overflow = dev_alloc<int>(1, &err);
  h->d_desc = dev_alloc<uint8_t>(O * 32, &err); h->d_out_desc = dev_alloc<uint8_t>(P * 32, &err);
  h->d_Ow = dev_alloc<double>(K * 3, &err); h->d_pos = dev_alloc<double>(P * 3, &err); h->d_normal = dev_alloc<double>(P * 3, &err);
  h->d_sf = dev_alloc<float>(CMOS_MAX_LEVELS, &err); h->d_min = dev_alloc<float>(P, &err); h->d_max = dev_alloc<float>(P, &err);
  if (err != cudaSuccess) {
    set_error("device allocation failed: %s", cudaGetErrorString(err));
    cmos_map_destroy(h);
    return CMOS_ERR_CUDA;
  }
  *out = h;
  return CMOS_OK;
}

static int check_csr(cmos_map* h, int n_points, const int32_t* obs_start, int* total) {
  CMOS_REQUIRE(n_points >= 0 && n_points <= h->p.max_points && obs_start, "n_points %d outside 0..%d", n_points, h->p.max_points);
  CMOS_REQUIRE(obs_start[0] == 0, "obs_start[0] must be 0");
  for (int p = 0; p < n_points; p++) CMOS_REQUIRE(obs_start[p + 1] >= obs_start[p], "obs_start must not decrease");
  *total = obs_start[n_points];
  CMOS_REQUIRE(*total <= h->p.max_observations, "%d observations exceed the handle's capacity %d", *total, h->p.max_observations);
  return CMOS_OK;
}

int cmos_map_distinctive_descriptors(cmos_map_t h, int32_t n_points, const int32_t* obs_start, const uint8_t* descriptors,
                                     int32_t* best_index, uint8_t* best_descriptor) {
  CMOS_REQUIRE(h && best_index, "null argument");
  int total = 0, rc;
  if ((rc = check_csr(h, n_points, obs_start, &total))) return rc;
  CMOS_REQUIRE(total == 0 || descriptors, "null argument");
  h->launches = 0;
  if (n_points == 0) return CMOS_OK;
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  CMOS_CUDA_OK(cudaMemcpyAsync(h->d_start, obs_start, (size_t)(n_points + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (total) CMOS_CUDA_OK(cudaMemcpyAsync(h->d_desc, descriptors, (size_t)total * 32, cudaMemcpyHostToDevice, st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_overflow, 0, sizeof(int), st));
  k_distinctive<<<(n_points + kMapWarps - 1) / kMapWarps, kMapWarps * 32, 0, st>>>(n_points, h->d_start, h->d_desc, h->d_best,
                                                                                 best_descriptor ? h->d_out_desc : nullptr,
                                                                                 h->d_overflow);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  int overflow = 0;
  CMOS_CUDA_OK(cudaMemcpyAsync(best_index, h->d_best, (size_t)n_points * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (best_descriptor)
    CMOS_CUDA_OK(cudaMemcpyAsync(best_descriptor, h->d_out_desc, (size_t)n_points * 32, cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(&overflow, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  if (overflow) {
    set_error("a map point has more than %d observations", kMapMaxObs);
    return CMOS_ERR_CAPACITY;
  }
  return CMOS_OK;
}

int cmos_map_update_normal_and_depth(cmos_map_t h, int32_t n_points, const int32_t* obs_start, const int32_t* obs_keyframe,
                                     int32_t n_keyframes, const double* camera_centers, const double* world_pos,
                                     const int32_t* ref_keyframe, const int32_t* ref_level, const float* scale_factors,
                                     int32_t n_levels, double* normal, float* min_distance, float* max_distance) {
  CMOS_REQUIRE(h && normal && min_distance && max_distance && scale_factors, "null argument");
  CMOS_REQUIRE(n_levels >= 1 && n_levels <= CMOS_MAX_LEVELS && n_keyframes >= 0 && n_keyframes <= h->p.max_keyframes, "bad sizes");
  int total = 0, rc;
  if ((rc = check_csr(h, n_points, obs_start, &total))) return rc;
  h->launches = 0;
  if (n_points == 0) return CMOS_OK;
  CMOS_REQUIRE(world_pos && ref_keyframe && ref_level && (total == 0 || (obs_keyframe && camera_centers)), "null argument");
  for (int i = 0; i < total; i++) CMOS_REQUIRE(obs_keyframe[i] >= 0 && obs_keyframe[i] < n_keyframes, "observation %d: keyframe out of range", i);
  for (int p = 0; p < n_points; p++)
    if (obs_start[p + 1] > obs_start[p])
      CMOS_REQUIRE(ref_keyframe[p] >= 0 && ref_keyframe[p] < n_keyframes && ref_level[p] >= 0 && ref_level[p] < n_levels,
                   "point %d: reference keyframe / level out of range", p);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  auto up = [&](void* dst, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
  CMOS_CUDA_OK(up(h->d_start, obs_start, (size_t)(n_points + 1) * sizeof(int)));
  CMOS_CUDA_OK(up(h->d_kf, obs_keyframe, (size_t)total * sizeof(int)));
  CMOS_CUDA_OK(up(h->d_Ow, camera_centers, (size_t)n_keyframes * 3 * sizeof(double)));
  CMOS_CUDA_OK(up(h->d_pos, world_pos, (size_t)n_points * 3 * sizeof(double)));
  CMOS_CUDA_OK(up(h->d_ref_kf, ref_keyframe, (size_t)n_points * sizeof(int)));
  CMOS_CUDA_OK(up(h->d_ref_level, ref_level, (size_t)n_points * sizeof(int)));
  CMOS_CUDA_OK(up(h->d_sf, scale_factors, (size_t)n_levels * sizeof(float)));
  CMOS_CUDA_OK(up(h->d_normal, normal, (size_t)n_points * 3 * sizeof(double)));      // points without observations keep theirs
  CMOS_CUDA_OK(up(h->d_min, min_distance, (size_t)n_points * sizeof(float)));
  CMOS_CUDA_OK(up(h->d_max, max_distance, (size_t)n_points * sizeof(float)));
  k_normal_depth<<<(n_points + 127) / 128, 128, 0, st>>>(n_points, h->d_start, h->d_kf, h->d_Ow, h->d_pos, h->d_ref_kf,
                                                        h->d_ref_level, h->d_sf, n_levels, h->d_normal, h->d_min, h->d_max);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  CMOS_CUDA_OK(cudaMemcpyAsync(normal, h->d_normal, (size_t)n_points * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(min_distance, h->d_min, (size_t)n_points * sizeof(float), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(max_distance, h->d_max, (size_t)n_points * sizeof(float), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_map_last_launch_count(cmos_map_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

}  // extern "C"
