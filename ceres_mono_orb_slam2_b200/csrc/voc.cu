// DBoW2 vocabulary transform for sm_100a: Frame::ComputeBoW / KeyFrame::ComputeBoW call
// orb_vocabulary_->transform(descriptors, bow_vector_, feature_vector_, 4) (src/Frame.cc:322-327, src/KeyFrame.cc:107-117;
// lib/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1260, FORB.cpp:81-101, BowVector.cpp:34-84, FeatureVector.cpp:31-45).
//
// The vocabulary (ORBvoc: k = 10, L = 6, 1.1 M nodes x 32 B = 35 MB) stays resident in HBM, flattened.
//  * k_voc_descend: one warp per descriptor walks the tree — lanes take the children of the current node, the warp
//    keeps the first minimum of the Hamming distance — and leaves word id, idf weight and the node at level L - levelsup.
//  * k_voc_assemble: one CTA turns the per-feature results into the two std::map-shaped outputs without atomics on
//    doubles: bitonic sort of (word, feature) and (node, feature) keys in shared memory, run heads by scan; a word's
//    value is its idf added once per occurrence (the reference accumulates `+= w`, not w * count), the L1 norm is
//    summed in ascending word order by one thread like BowVector::normalize does, then every value is divided.
// Compiled with --fmad=false.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "cmos_common.h"

namespace cmos {

constexpr int kVocMaxFeatures = 8192;     // features per transform call (keys sorted in shared memory)
constexpr int kVocThreads = 1024;

__global__ void __launch_bounds__(256) k_voc_descend(const int* __restrict__ child_start, const int* __restrict__ children,
                                                     const uint8_t* __restrict__ node_desc,
                                                     const double* __restrict__ node_weight,
                                                     const int* __restrict__ node_word, int nid_level,
                                                     const uint8_t* __restrict__ features, int n, int* __restrict__ f_word,
                                                     int* __restrict__ f_node, double* __restrict__ f_weight) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  uint32_t f[8];
#pragma unroll
  for (int w = 0; w < 8; w++) f[w] = __ldg((const uint32_t*)(features + 32 * (size_t)i) + w);
  int node = 0, level = 0, nid = 0;
  for (;;) {
    const int c0 = child_start[node], c1 = child_start[node + 1];
    if (c1 <= c0) break;                                  // leaf
    ++level;
    unsigned best = 0xffffffffu;                          // (distance << 16 | position): first minimum wins
    for (int c = c0 + lane; c < c0 + ((c1 - c0 + 31) & ~31); c += 32) {
      unsigned key = 0xffffffffu;
      if (c < c1) {
        const uint32_t* dj = (const uint32_t*)(node_desc + 32 * (size_t)children[c]);
        int d = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) d += __popc(f[w] ^ __ldg(dj + w));
        key = ((unsigned)d << 16) | (unsigned)(c - c0);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
      best = min(best, key);
    }
    node = children[c0 + (int)(best & 0xffffu)];
    if (level == nid_level) nid = node;
  }
  if (lane == 0) { f_word[i] = node_word[node]; f_node[i] = nid; f_weight[i] = node_weight[node]; }
}

__device__ void bitonic_sort(unsigned long long* key, int np2) {
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < np2; t += blockDim.x) {
        const int x = t ^ j;
        if (x > t) {
          const unsigned long long a = key[t], b = key[x];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { key[t] = b; key[x] = a; }
        }
      }
      __syncthreads();
    }
}

// out_count[0] = number of words, out_count[1] = number of feature-vector nodes
__global__ void __launch_bounds__(kVocThreads) k_voc_assemble(int n, const int* __restrict__ f_word,
                                                              const int* __restrict__ f_node,
                                                              const double* __restrict__ f_weight, int* __restrict__ bow_words,
                                                              double* __restrict__ bow_values, int* __restrict__ fv_nodes,
                                                              int* __restrict__ fv_start, int* __restrict__ fv_features,
                                                              int* __restrict__ out_count) {
  extern __shared__ unsigned long long s_key[];          // [np2]
  __shared__ int s_total, s_scan[kVocThreads];
  __shared__ double s_norm;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const int tid = threadIdx.x;
  for (int pass = 0; pass < 2; pass++) {
    const int* id = pass == 0 ? f_word : f_node;
    for (int t = tid; t < np2; t += kVocThreads)
      s_key[t] = (t < n && f_weight[t] > 0) ? (((unsigned long long)(unsigned)id[t] << 32) | (unsigned)t) : ~0ull;   // w > 0: not stopped
    __syncthreads();
    bitonic_sort(s_key, np2);
    // run heads -> positions (block scan over per-thread counts of contiguous chunks)
    const int per = (np2 + kVocThreads - 1) / kVocThreads, lo = tid * per, hi = min(lo + per, np2);
    int cnt = 0;
    for (int t = lo; t < hi; t++)
      if (s_key[t] != ~0ull && (t == 0 || (s_key[t] >> 32) != (s_key[t - 1] >> 32))) cnt++;
    s_scan[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < kVocThreads; o <<= 1) {
      const int v = tid >= o ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    int pos = s_scan[tid] - cnt;
    if (tid == kVocThreads - 1) s_total = s_scan[tid];
    for (int t = lo; t < hi; t++) {
      const unsigned long long k = s_key[t];
      if (k == ~0ull) continue;
      const bool head = t == 0 || (k >> 32) != (s_key[t - 1] >> 32);
      if (pass == 0) {
        if (head) {
          // value = idf added once per occurrence, in feature order (BowVector::addWeight)
          const double w = f_weight[(unsigned)k];
          double v = w;
          for (int u = t + 1; u < np2 && (s_key[u] >> 32) == (k >> 32); u++) v += w;
          bow_words[pos] = (int)(k >> 32);
          bow_values[pos] = v;
          pos++;
        }
      } else {
        if (head) { fv_nodes[pos] = (int)(k >> 32); fv_start[pos] = t; pos++; }
        fv_features[t] = (int)(unsigned)k;               // valid keys sort to the front: position t is final
      }
    }
    __syncthreads();
    const int total = s_total;
    if (pass == 0) {
      if (tid == 0) {                                    // BowVector::normalize(L1): ascending word order
        double norm = 0.0;
        for (int u = 0; u < total; u++) norm += fabs(bow_values[u]);
        s_norm = norm;
        out_count[0] = total;
      }
      __syncthreads();
      const double norm = s_norm;
      if (norm > 0.0)
        for (int u = tid; u < total; u += kVocThreads) bow_values[u] /= norm;
    } else if (tid == 0) {
      int valid = 0;                                     // number of non-stopped features = end of the last run
      int a = 0, b = np2;                                // first index holding the sentinel
      while (a < b) { const int m = (a + b) >> 1; if (s_key[m] == ~0ull) b = m; else a = m + 1; }
      valid = a;
      fv_start[total] = valid;
      out_count[1] = total;
    }
    __syncthreads();
  }
}

}  // namespace cmos

using namespace cmos;

struct cmos_voc {
  int device = 0, n_nodes = 0, L = 0, launches = 0;
  cudaStream_t stream = nullptr;
  int *d_child_start = nullptr, *d_children = nullptr, *d_word = nullptr;
  uint8_t* d_desc = nullptr;
  double* d_weight = nullptr;
  // per-call buffers
  uint8_t* d_feat = nullptr;
  int *d_fword = nullptr, *d_fnode = nullptr, *d_bw = nullptr, *d_fn = nullptr, *d_fs = nullptr, *d_ff = nullptr, *d_cnt = nullptr;
  double *d_fweight = nullptr, *d_bv = nullptr;
};

extern "C" {

int cmos_voc_destroy(cmos_voc_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->device);
  for (void* b : {(void*)h->d_child_start, (void*)h->d_children, (void*)h->d_word, (void*)h->d_desc, (void*)h->d_weight,
                  (void*)h->d_feat, (void*)h->d_fword, (void*)h->d_fnode, (void*)h->d_bw, (void*)h->d_fn, (void*)h->d_fs,
                  (void*)h->d_ff, (void*)h->d_cnt, (void*)h->d_fweight, (void*)h->d_bv})
    if (b) cudaFree(b);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CMOS_OK;
}

int cmos_voc_create(int32_t n_nodes, const int32_t* child_start, const int32_t* children, const uint8_t* node_descriptors,
                    const double* node_weights, const int32_t* node_word_ids, int32_t depth_levels, int32_t device,
                    cmos_voc_t* out) {
  CMOS_REQUIRE(out && child_start && children && node_descriptors && node_weights && node_word_ids, "null argument");
  CMOS_REQUIRE(n_nodes >= 2 && depth_levels >= 1 && child_start[0] == 0, "bad vocabulary");
  const int n_children = child_start[n_nodes];
  CMOS_REQUIRE(n_children == n_nodes - 1, "a tree with %d nodes has %d child links, not %d", n_nodes, n_nodes - 1, n_children);
  for (int i = 0; i < n_nodes; i++) {
    CMOS_REQUIRE(child_start[i + 1] >= child_start[i] && child_start[i + 1] - child_start[i] <= 65535, "bad child range at node %d", i);
    for (int c = child_start[i]; c < child_start[i + 1]; c++)
      CMOS_REQUIRE(children[c] > 0 && children[c] < n_nodes, "child id out of range at node %d", i);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
  CMOS_CUDA_OK(cudaSetDevice(device));
  cmos_voc* h = new cmos_voc();
  h->device = device; h->n_nodes = n_nodes; h->L = depth_levels;
  cudaError_t err = cudaSuccess;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) err = cudaErrorUnknown;
  const size_t N = n_nodes, F = kVocMaxFeatures;
  h->d_child_start = dev_alloc<int>(N + 1, &err); h->d_children = dev_alloc<int>(N, &err); h->d_word = dev_alloc<int>(N, &err);
  h->d_desc = dev_alloc<uint8_t>(N * 32, &err); h->d_weight = dev_alloc<double>(N, &err);
  h->d_feat = dev_alloc<uint8_t>(F * 32, &err); h->d_fword = dev_alloc<int>(F, &err); h->d_fnode = dev_alloc<int>(F, &err);
  h->d_bw = dev_alloc<int>(F, &err); h->d_fn = dev_alloc<int>(F, &err); h->d_fs = dev_alloc<int>(F + 1, &err);
  h->d_ff = dev_alloc<int>(F, &err); h->d_cnt = dev_alloc<int>(2, &err);
  h->d_fweight = dev_alloc<double>(F, &err); h->d_bv = dev_alloc<double>(F, &err);
  if (err != cudaSuccess) {
    set_error("device allocation failed: %s", cudaGetErrorString(err));
    cmos_voc_destroy(h);
    return CMOS_ERR_CUDA;
  }
  cudaMemcpy(h->d_child_start, child_start, (N + 1) * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_children, children, (size_t)n_children * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_word, node_word_ids, N * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_desc, node_descriptors, N * 32, cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_weight, node_weights, N * sizeof(double), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_voc_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, kVocMaxFeatures * 8);
  CMOS_CUDA_OK(cudaGetLastError());
  *out = h;
  return CMOS_OK;
}

int cmos_voc_transform(cmos_voc_t h, const uint8_t* descriptors, int32_t n, int32_t levelsup, int32_t* bow_words,
                       double* bow_values, int32_t* n_words, int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features,
                       int32_t* n_fv_nodes) {
  CMOS_REQUIRE(h && bow_words && bow_values && n_words && fv_nodes && fv_start && fv_features && n_fv_nodes, "null argument");
  CMOS_REQUIRE(n >= 0 && n <= kVocMaxFeatures && (n == 0 || descriptors), "n %d outside 0..%d", n, kVocMaxFeatures);
  *n_words = 0; *n_fv_nodes = 0; fv_start[0] = 0;
  h->launches = 0;
  if (n == 0) return CMOS_OK;
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  CMOS_CUDA_OK(cudaMemcpyAsync(h->d_feat, descriptors, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  k_voc_descend<<<(n + 7) / 8, 256, 0, st>>>(h->d_child_start, h->d_children, h->d_desc, h->d_weight, h->d_word,
                                            h->L - levelsup, h->d_feat, n, h->d_fword, h->d_fnode, h->d_fweight);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  k_voc_assemble<<<1, kVocThreads, (size_t)np2 * 8, st>>>(n, h->d_fword, h->d_fnode, h->d_fweight, h->d_bw, h->d_bv, h->d_fn,
                                                         h->d_fs, h->d_ff, h->d_cnt);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 2;
  int cnt[2] = {0, 0};
  CMOS_CUDA_OK(cudaMemcpyAsync(cnt, h->d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  *n_words = cnt[0]; *n_fv_nodes = cnt[1];
  CMOS_CUDA_OK(cudaMemcpyAsync(bow_words, h->d_bw, (size_t)cnt[0] * sizeof(int), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(bow_values, h->d_bv, (size_t)cnt[0] * sizeof(double), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(fv_nodes, h->d_fn, (size_t)cnt[1] * sizeof(int), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(fv_start, h->d_fs, (size_t)(cnt[1] + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  const int n_valid = fv_start[cnt[1]];
  if (n_valid) CMOS_CUDA_OK(cudaMemcpy(fv_features, h->d_ff, (size_t)n_valid * sizeof(int), cudaMemcpyDeviceToHost));
  return CMOS_OK;
}

int cmos_voc_last_launch_count(cmos_voc_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

}  // extern "C"
