// ORB matcher for sm_100a: Frame grid (CSR), SearchByProjection(frame, last frame) and
// SearchByProjection(frame, map points), isInFrustum — batched over frames, one CTA per frame.
//
// Behaviour follows src/ORBmatcher.cc:42-126,1161-1271,1386-1437 and src/Frame.cc:158-173,191-320.
// The reference's searches are greedy and order dependent (a keypoint claimed by an earlier map point is
// skipped by later ones).  Here the Hamming work is done in parallel — every warp of the CTA takes
// queries, walks the grid window in the reference's iteration order and leaves a short, ordered
// candidate list per query in shared memory — and then ONE warp replays the queries in order against the
// `claimed` bytes, which reproduces the sequential result exactly.  Queries whose list overflowed are
// re-enumerated by that warp, so the result never depends on the list capacity.
//
// Compiled with --fmad=false (projection and window arithmetic must round like the unfused CPU floats).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cmos_common.h"
#include "match_device.cuh"

namespace cmos {

constexpr int kSearchThreads = 1024;
constexpr int kListCap = 8;        // candidates kept per query for the sequential replay (overflow -> re-enumerate)
constexpr int kListCapPts = 8;

// -------------------------------------------------------------------------------------------------
// Frame::AssignFeaturesToGrid (Frame.cc:158-173) + PosInGrid (:309-320) as CSR, one CTA per frame.
__global__ void __launch_bounds__(256) k_build_grid(cmos_camera cam, const cmos_keypoint* __restrict__ kps,
                                                    const int* __restrict__ counts, int stride, int max_kp,
                                                    int* __restrict__ grid_start, int* __restrict__ grid_idx,
                                                    float4* __restrict__ cells) {
  extern __shared__ int sm[];
  int* cnt = sm;                              // [kCells + 1]
  short* cell_of = (short*)(cnt + kCells + 1);   // [n]
  __shared__ int warp_tmp[8];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(counts[f], stride);
  const cmos_keypoint* kp = kps + (long long)f * stride;
  for (int i = tid; i <= kCells; i += 256) cnt[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    const int px = (int)roundf((kp[i].x - cam.min_x) * cam.grid_element_width_inv);
    const int py = (int)roundf((kp[i].y - cam.min_y) * cam.grid_element_height_inv);
    int c = -1;
    if (px >= 0 && px < kCols && py >= 0 && py < kRows) { c = px * kRows + py; atomicAdd(&cnt[c], 1); }
    cell_of[i] = (short)c;
  }
  __syncthreads();
  // exclusive scan of 3072 counts: 12 per thread
  constexpr int per = kCells / 256;
  int local[per], sum = 0;
#pragma unroll
  for (int j = 0; j < per; j++) { local[j] = cnt[tid * per + j]; sum += local[j]; }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) warp_tmp[tid >> 5] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < (tid >> 5); w++) base += warp_tmp[w];
  int run = base + incl - sum;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < per; j++) { cnt[tid * per + j] = run; run += local[j]; }
  if (tid == 255) cnt[kCells] = run;
  __syncthreads();
  int* gs = grid_start + (long long)f * (kCells + 1);
  for (int i = tid; i <= kCells; i += 256) gs[i] = cnt[i];
  __syncthreads();   // cnt[] becomes the per-cell cursor below
  // Stable scatter (insertion order inside a cell, Frame.cc:166-172): warp 0 walks the keypoints 32 at a time in
  // index order; lanes that share a cell take consecutive slots from that cell's cursor.
  int* gi = grid_idx + (long long)f * max_kp;
  if (tid < 32) {
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + tid;
      const int c = i < n ? (int)cell_of[i] : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, c);
      const int leader = __ffs(peers) - 1;
      int cur = 0;
      if (c >= 0 && tid == leader) { cur = cnt[c]; cnt[c] = cur + __popc(peers); }
      cur = __shfl_sync(0xffffffffu, cur, leader);
      if (c >= 0) gi[cur + __popc(peers & ((1u << tid) - 1))] = i;
      __syncwarp();
    }
  }
  if (cells) {          // the keypoints in cell order, by every thread (octave < 128, i < 2^24: FrameDev::cells)
    __syncthreads();
    const int n_in = cnt[kCells];                   // keypoints inside the grid (the scan.s total; the cursors are cnt[0..kCells-1])
    float4* cf = cells + (long long)f * max_kp;
    for (int pos = tid; pos < n_in; pos += 256) {
      const int i = gi[pos];
      cf[pos] = make_float4(kp[i].x, kp[i].y, __int_as_float((kp[i].octave << 24) | i), 0.f);
    }
  }
}

int launch_build_grid(const cmos_camera& cam, const cmos_keypoint* kps, const int* counts, int stride, int max_kp,
                      int* grid_start, int* grid_idx, int n_frames, cudaStream_t st, float4* cells) {
  const size_t smem = (kCells + 1) * sizeof(int) + (size_t)stride * sizeof(short) + 16;
  k_build_grid<<<n_frames, 256, smem, st>>>(cam, kps, counts, stride, max_kp, grid_start, grid_idx, cells);
  CMOS_CUDA_OK(cudaGetLastError());
  return CMOS_OK;
}

// -------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th), one CTA per frame pair.
struct SearchFrameArgs {
  const double* Tcw;                 // [B][16]
  const cmos_keypoint* last_kps;     // [B][last_stride]
  const int* last_counts;
  const uint8_t* last_flags;
  const double* last_xw;
  const uint8_t* last_desc;
  int last_stride;
  float th;
  int check_ori;
  uint8_t* claimed;                  // [B][stride] or null
  int* match;                        // [B][stride]
  int* nmatches;                     // [B]
  const float4* cells;               // [B][max_kp] keypoints in cell order (FrameDev::cells), or null
};

struct QueryFrame { float u, v, radius; int oct; bool ok; };

__device__ __forceinline__ QueryFrame project_last(const cmos_camera& cam, const double* __restrict__ T,
                                                   const double* __restrict__ X, int oct) {
  QueryFrame q;
  q.ok = false;
  q.oct = oct;
  const double pcx = (T[0] * X[0] + T[1] * X[1]) + T[2] * X[2] + T[3];
  const double pcy = (T[4] * X[0] + T[5] * X[1]) + T[6] * X[2] + T[7];
  const double pcz = (T[8] * X[0] + T[9] * X[1]) + T[10] * X[2] + T[11];
  const float xc = (float)pcx, yc = (float)pcy;
  const float invzc = (float)(1.0 / pcz);
  if (invzc < 0) return q;
  q.u = cam.fx * xc * invzc + cam.cx;
  q.v = cam.fy * yc * invzc + cam.cy;
  if (q.u < cam.min_x || q.u > cam.max_x) return q;
  if (q.v < cam.min_y || q.v > cam.max_y) return q;
  q.radius = 0.f;
  q.ok = true;
  return q;
}

// Phase 1 — one warp per last-frame keypoint (query), the whole batch in one launch: project, walk the grid
// window in the reference's candidate order, Hamming against every candidate; leaves in global memory the
// ordered list of candidates with distance <= TH_HIGH (first kListCap), their count, and the first minimum.
constexpr int kSfWarps = 8;
constexpr uint32_t kNoBest = 0x7fffffffu;   // low 31 bits of a best word when the query has no candidate

#ifndef CMOS_SF_MINBLOCKS
#define CMOS_SF_MINBLOCKS 8      // 32 registers, 64 warps per SM: search 0.247 -> 0.229 ms per 64 frames (6: 0.243; unbounded 48 registers: 0.247)
#endif
#if CMOS_SF_MINBLOCKS > 0
__global__ void __launch_bounds__(kSfWarps * 32, CMOS_SF_MINBLOCKS) k_sf_lists(
#else
__global__ void __launch_bounds__(kSfWarps * 32) k_sf_lists(
#endif
    cmos_camera cam, const cmos_keypoint* __restrict__ kps,
                                                           const uint8_t* __restrict__ desc,
                                                           const int* __restrict__ counts, int stride,
                                                           const int* __restrict__ grid_start,
                                                           const int* __restrict__ grid_idx, int max_kp,
                                                           SearchFrameArgs a, uint32_t* __restrict__ g_lists,
                                                           uint32_t* __restrict__ g_best, uint16_t* __restrict__ g_cnt) {
  const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * kSfWarps + warp;
  const int nq = min(a.last_counts[f], a.last_stride);
  // The projection and the cell window of a query are scalar work (fp64 pose product, a division, four floor / ceil): done by
  // its warp they cost a full warp instruction each, ~190 per query.  Thread t of the CTA does them for query t of the CTA's
  // eight instead — one warp's worth of issue slots for all eight — and hands {u, v, radius, window} over in shared memory.
  __shared__ float4 s_uvr[kSfWarps];     // u, v, radius, (min_cx | max_cx << 8 | min_cy << 16 | max_cy << 24) or -1: no window
  if (threadIdx.x < kSfWarps) {
    const int qq = blockIdx.x * kSfWarps + threadIdx.x;
    float4 o = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (qq < nq && (a.last_flags[(long long)f * a.last_stride + qq] & 1)) {
      const int oct = a.last_kps[(long long)f * a.last_stride + qq].octave;
      const QueryFrame Q = project_last(cam, a.Tcw + (long long)f * 16, a.last_xw + ((long long)f * a.last_stride + qq) * 3, oct);
      if (Q.ok) {
        const float radius = a.th * cam.scale_factors[oct];
        const Window w = make_window(cam, Q.u, Q.v, radius);
        if (w.ok) o = make_float4(Q.u, Q.v, radius, __int_as_float(w.min_cx | (w.max_cx << 8) | (w.min_cy << 16) | (w.max_cy << 24)));
      }
    }
    s_uvr[threadIdx.x] = o;
  }
  __syncthreads();
  if (q >= nq) return;
  const int n = min(counts[f], stride);
  FrameDev F{kps + (long long)f * stride, desc + (long long)f * stride * 32,
             grid_start + (long long)f * (kCells + 1), grid_idx + (long long)f * max_kp, n,
             a.cells ? a.cells + (long long)f * max_kp : nullptr};
  const cmos_keypoint* last = a.last_kps + (long long)f * a.last_stride;
  const uint8_t flag = a.last_flags[(long long)f * a.last_stride + q];
  const size_t slot = (size_t)f * a.last_stride + q;
  // lanes 0..kListCap-1 keep the kListCap smallest (distance, candidate order) keys seen so far, ascending
  uint32_t tkey = 0xffffffffu;
  int tidx = 0, total = 0;
  if (flag & 1) {
    const int oct = last[q].octave;
    const float4 uvr = s_uvr[warp];
    const int wbits = __float_as_int(uvr.w);
    {
      struct { float u, v; } Q{uvr.x, uvr.y};
      const float radius = uvr.z;
      const Window w{wbits & 0xff, (wbits >> 8) & 0xff, (wbits >> 16) & 0xff, (wbits >> 24) & 0xff, wbits >= 0};
      if (w.ok) {
        uint32_t dq[8];
        const uint32_t* dsrc = (const uint32_t*)(a.last_desc + ((long long)f * a.last_stride + q) * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) dq[i] = __ldg(dsrc + i);
        walk_window(F, w, Q.u, Q.v, radius, oct - 1, oct + 1, lane, [&](int idx, bool pass) {
          int d = 256;
          if (pass) d = hamming256(dq, F.desc + 32 * (size_t)idx);
          const bool keep = pass && d <= CMOS_TH_HIGH;
          const unsigned m = __ballot_sync(0xffffffffu, keep);
          if (m) {
            const int pos = total + __popc(m & ((1u << lane) - 1));
            total += __popc(m);
            uint32_t nkey = keep ? (((uint32_t)d << 16) | (uint32_t)min(pos, 0xffff)) : 0xffffffffu;
            uint32_t okey = 0xffffffffu;
            int oidx = 0;
#pragma unroll
            for (int r = 0; r < kListCap; r++) {
              const uint32_t mine = min(nkey, tkey);
              uint32_t wm = mine;
#pragma unroll
              for (int o = 16; o; o >>= 1) wm = min(wm, __shfl_xor_sync(0xffffffffu, wm, o));
              if (wm == 0xffffffffu) break;
              const int owner = __ffs(__ballot_sync(0xffffffffu, mine == wm)) - 1;
              const int from_new = __shfl_sync(0xffffffffu, (int)(nkey == wm), owner);
              const int sel = __shfl_sync(0xffffffffu, from_new ? idx : tidx, owner);
              if (lane == owner) { if (from_new) nkey = 0xffffffffu; else tkey = 0xffffffffu; }
              if (lane == r) { okey = wm; oidx = sel; }
            }
            tkey = okey; tidx = oidx;
          }
        });
      }
    }
  }
  if (lane < kListCap && lane < total) g_lists[slot * kListCap + lane] = ((tkey >> 16) << 16) | (uint32_t)tidx;
  if (lane == 0) {
    g_cnt[slot] = (uint16_t)min(total, 65535);
    g_best[slot] = (total > 0 ? (((tkey >> 16) << 16) | (uint32_t)tidx) : kNoBest) | ((uint32_t)((flag >> 1) & 1) << 31);
  }
}

// Phase 2 — one CTA per frame pair resolves the greedy, order-dependent assignment (ORBmatcher.cc:1176-1250).
// Sequentially, query q takes the first entry of its (distance, order)-sorted list that no CLAIMING query j < q
// holds (a claiming query is one whose map point has Observations() > 0, :1220-1221).  That is a fixed point and
// is computed in parallel: every query points at a list entry; owner[idx] = lowest claiming query that ever
// pointed at idx (atomicMin — a query only leaves an entry when a lower one owns it, so "ever" == "currently");
// queries whose entry is owned by a lower query advance; repeat until nothing moves.  match[idx] is the highest
// query that ends on idx (the last writer), every final pick is one rotation-histogram event and one nmatches++.
// A query that exhausts a truncated list (more than kListCap candidates, all taken) re-walks its window for the best
// keypoint no lower query holds, so the result never depends on the list capacity.  After kMaxRounds rounds
// (pathological chains) the frame is replayed sequentially by one warp instead — the same answer, slowly.
constexpr int kReplayThreads = 256;

__global__ void __launch_bounds__(kReplayThreads) k_sf_replay(cmos_camera cam, const cmos_keypoint* __restrict__ kps,
                                                             const uint8_t* __restrict__ desc,
                                                             const int* __restrict__ counts, int stride,
                                                             const int* __restrict__ grid_start,
                                                             const int* __restrict__ grid_idx, int max_kp,
                                                             SearchFrameArgs a, const uint32_t* __restrict__ g_lists,
                                                             const uint32_t* __restrict__ g_best,
                                                             const uint16_t* __restrict__ g_cnt, int stage_lists) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(counts[f], stride);
  const int nq = min(a.last_counts[f], a.last_stride);
  // carve: lists[nq][kListCap] u32 | best[nq] u32 | match[n] i32 | ev_idx[nq] u16 | ev_q[nq] u16 | cnt[nq] u16 |
  //        claimed[n] u8 | ev_bin[nq] u8
  // (large keypoint counts: the lists stay in global memory / L2 and only the small per-query state is staged)
  const uint32_t* lists = stage_lists ? (const uint32_t*)smem_raw : g_lists + (size_t)f * a.last_stride * kListCap;
  uint32_t* s_best = (uint32_t*)smem_raw + (stage_lists ? (size_t)a.last_stride * kListCap : 0);
  int* s_match = (int*)(s_best + a.last_stride);
  uint16_t* ev_idx = (uint16_t*)(s_match + stride);
  uint16_t* ev_q = ev_idx + a.last_stride;
  uint16_t* cnt = ev_q + a.last_stride;
  uint8_t* s_claimed = (uint8_t*)(cnt + a.last_stride);
  uint8_t* ev_bin = s_claimed + stride;
  uint8_t* ptr = ev_bin + a.last_stride;                                   // [nq] current list position, 0xff = none
  int* owner = (int*)(((uintptr_t)(ptr + a.last_stride) + 3) & ~(uintptr_t)3);   // [n]
  uint16_t* ext = (uint16_t*)(owner + stride);                             // [nq] pick found by a window re-walk
  __shared__ int s_nmatch, s_nev, s_hist[CMOS_HISTO_LENGTH], s_keep[3], s_changed, s_nover;

  FrameDev F{kps + (long long)f * stride, desc + (long long)f * stride * 32,
             grid_start + (long long)f * (kCells + 1), grid_idx + (long long)f * max_kp, n,
             a.cells ? a.cells + (long long)f * max_kp : nullptr};
  const cmos_keypoint* last = a.last_kps + (long long)f * a.last_stride;
  const double* lxw = a.last_xw + (long long)f * a.last_stride * 3;
  const uint8_t* ldesc = a.last_desc + (long long)f * a.last_stride * 32;
  const double* T = a.Tcw + (long long)f * 16;
  const size_t base = (size_t)f * a.last_stride;

  if (stage_lists) {
    const uint4* src = (const uint4*)(g_lists + base * kListCap);
    uint4* dst = (uint4*)smem_raw;
    for (int i = tid; i < nq * kListCap / 4; i += kReplayThreads) dst[i] = src[i];
  }
  for (int i = tid; i < nq; i += kReplayThreads) { s_best[i] = g_best[base + i]; cnt[i] = g_cnt[base + i]; }
  for (int i = tid; i < n; i += kReplayThreads) {
    s_match[i] = -1;
    s_claimed[i] = a.claimed ? a.claimed[(long long)f * stride + i] : 0;
  }
  if (tid < CMOS_HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) { s_nmatch = 0; s_nev = 0; }
  __syncthreads();

  // ---- parallel fixed point ----
  // ptr[q]: 0..kListCap-1 = position in the stored list, kExt = pick found by re-walking the window (ext[q]),
  // kNone = no candidate left (final: the set of blocked keypoints only grows).
  constexpr int kExt = 0xfe, kNone = 0xff, kMaxRounds = 64;
  for (int i = tid; i < n; i += kReplayThreads) owner[i] = s_claimed[i] ? -1 : 0x7fffffff;
  for (int q = tid; q < nq; q += kReplayThreads) ptr[q] = cnt[q] > 0 ? 0 : kNone;
  uint16_t* over = ev_q;            // queries that must re-walk their window this round (ev_q is free until the end)
  bool sequential = false;
  for (int round = 0;; round++) {
    __syncthreads();
    if (tid == 0) { s_changed = 0; s_nover = 0; }
    // A: every claiming query stamps the keypoint it points at
    for (int q = tid; q < nq; q += kReplayThreads) {
      const int k = ptr[q];
      if (k == kNone || !(s_best[q] >> 31)) continue;
      atomicMin(&owner[k == kExt ? (int)ext[q] : (int)(lists[(size_t)q * kListCap + k] & 0xffff)], q);
    }
    __syncthreads();
    // B: queries whose keypoint belongs to a lower query move on
    for (int q = tid; q < nq; q += kReplayThreads) {
      int k = ptr[q];
      if (k == kNone) continue;
      if (k == kExt) {
        if (owner[ext[q]] < q) over[atomicAdd(&s_nover, 1)] = (uint16_t)q;
        continue;
      }
      const int c = cnt[q], lim = min(c, kListCap), k0 = k;
      while (k < lim && owner[lists[(size_t)q * kListCap + k] & 0xffff] < q) k++;
      if (k == lim) {
        if (c > kListCap) { over[atomicAdd(&s_nover, 1)] = (uint16_t)q; continue; }   // truncated list exhausted
        k = kNone;
      }
      if (k != k0) { ptr[q] = (uint8_t)k; s_changed = 1; }
    }
    __syncthreads();
    // C: one warp per exhausted query walks its window again: best keypoint not held by a lower query
    const int nover = s_nover;
    for (int i = warp; i < nover; i += kReplayThreads / 32) {
      const int q = over[i];
      const int oct = last[q].octave;
      QueryFrame Q = project_last(cam, T, lxw + 3 * q, oct);
      const float radius = a.th * cam.scale_factors[oct];
      const Window w = make_window(cam, Q.u, Q.v, radius);
      uint32_t dq[8];
#pragma unroll
      for (int j = 0; j < 8; j++) dq[j] = __ldg((const uint32_t*)(ldesc + 32 * (size_t)q) + j);
      int bd = 256, bi = -1;
      walk_window(F, w, Q.u, Q.v, radius, oct - 1, oct + 1, lane, [&](int idx, bool pass) {
        int d = 256;
        if (pass && owner[idx] >= q) d = hamming256(dq, F.desc + 32 * (size_t)idx);
        unsigned key = ((unsigned)d << 8) | (unsigned)lane, mk = key;
#pragma unroll
        for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
        const int cd = (int)(mk >> 8);
        if (cd < bd) { bd = cd; bi = __shfl_sync(0xffffffffu, idx, mk & 31); }
      });
      if (lane == 0) {
        if (bd <= CMOS_TH_HIGH && bi >= 0) { ptr[q] = kExt; ext[q] = (uint16_t)bi; }
        else ptr[q] = kNone;
        s_changed = 1;
      }
    }
    __syncthreads();
    if (!s_changed) break;
    if (round >= kMaxRounds) { sequential = true; break; }   // pathological chains: exact sequential replay instead
  }
  if (!sequential) {
    for (int q = tid; q < nq; q += kReplayThreads) {
      const int k = ptr[q];
      if (k == kNone) continue;
      const int idx = k == kExt ? (int)ext[q] : (int)(lists[(size_t)q * kListCap + k] & 0xffff);
      atomicMax(&s_match[idx], q);
      if (s_best[q] >> 31) s_claimed[idx] = 1;
    }
    __syncthreads();     // `over` aliases ev_q: events are written only after the last round has finished with it
    for (int q = tid; q < nq; q += kReplayThreads) {
      const int k = ptr[q];
      if (k == kNone) continue;
      const int idx = k == kExt ? (int)ext[q] : (int)(lists[(size_t)q * kListCap + k] & 0xffff);
      const int e = atomicAdd(&s_nev, 1);
      ev_idx[e] = (uint16_t)idx; ev_q[e] = (uint16_t)q;
    }
    __syncthreads();
    if (tid == 0) s_nmatch = s_nev;
  }

  if (sequential && warp == 0) {
    int q = 0, nev = 0;
    for (;;) {
      int stop = nq;
      if (lane == 0) {
        // lists are sorted by (distance, candidate order): the first unclaimed entry is the reference's answer
        while (q < nq) {
          const int c = cnt[q];
          if (c > 0) {
            const uint32_t claim = s_best[q] >> 31;
            const int lim = min(c, kListCap);
            int k = 0;
            for (; k < lim; k++) {
              const int idx = lists[(size_t)q * kListCap + k] & 0xffff;
              if (!s_claimed[idx]) {
                s_match[idx] = q;
                s_claimed[idx] = (uint8_t)claim;
                ev_idx[nev] = (uint16_t)idx; ev_q[nev] = (uint16_t)q; nev++;
                break;
              }
            }
            if (k == lim && c > kListCap) break;   // every listed candidate is claimed and there are more: walk again
          }
          q++;
        }
        stop = q;
      }
      q = __shfl_sync(0xffffffffu, stop, 0);
      if (q >= nq) break;
      nev = __shfl_sync(0xffffffffu, nev, 0);
      __syncwarp();
      {
        const int oct = last[q].octave;
        QueryFrame Q = project_last(cam, T, lxw + 3 * q, oct);
        const float radius = a.th * cam.scale_factors[oct];
        const Window w = make_window(cam, Q.u, Q.v, radius);
        uint32_t dq[8];
#pragma unroll
        for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(ldesc + 32 * (size_t)q) + i);
        int bd = 256, bi = -1;
        walk_window(F, w, Q.u, Q.v, radius, oct - 1, oct + 1, lane, [&](int idx, bool pass) {
          int d = 256;
          if (pass && !s_claimed[idx]) d = hamming256(dq, F.desc + 32 * (size_t)idx);
          unsigned key = ((unsigned)d << 8) | (unsigned)lane, mk = key;
#pragma unroll
          for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
          const int cd = (int)(mk >> 8);
          if (cd < bd) { bd = cd; bi = __shfl_sync(0xffffffffu, idx, mk & 31); }
        });
        if (bd <= CMOS_TH_HIGH && bi >= 0 && lane == 0) {
          s_match[bi] = q;
          s_claimed[bi] = (uint8_t)(s_best[q] >> 31);
          ev_idx[nev] = (uint16_t)bi; ev_q[nev] = (uint16_t)q;
        }
        if (bd <= CMOS_TH_HIGH && bi >= 0) nev++;
      }
      __syncwarp();
      q++;
    }
    if (lane == 0) { s_nev = nev; s_nmatch = nev; }
  }
  __syncthreads();
  const int nev = s_nev;
  if (a.check_ori) {
    const float factor = 1.0f / CMOS_HISTO_LENGTH;
    for (int e = tid; e < nev; e += kReplayThreads) {
      float rot = last[ev_q[e]].angle - F.kps[ev_idx[e]].angle;
      if (rot < 0.0f) rot += 360.0f;
      int bin = (int)roundf(rot * factor);
      if (bin == CMOS_HISTO_LENGTH) bin = 0;
      ev_bin[e] = (uint8_t)bin;
      atomicAdd(&s_hist[bin], 1);
    }
    __syncthreads();
    // ComputeThreeMaxima (ORBmatcher.cc:1386-1418)
    if (tid == 0) {
      int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
      for (int i = 0; i < CMOS_HISTO_LENGTH; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
        else if (s > max3) { max3 = s; i3 = i; }
      }
      if (max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
      else if (max3 < 0.1f * (float)max1) { i3 = -1; }
      s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3;
    }
    __syncthreads();
    int removed = 0;
    for (int e = tid; e < nev; e += kReplayThreads) {
      const int b = ev_bin[e];
      // ORBmatcher.cc:1262 nulls map_points_[idx]: the keypoint is no longer claimed either (a 2 x th retry, Tracking.cc:636-639,
      // or a later SearchLocalPoints on the same frame may match it again)
      if (b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) { s_match[ev_idx[e]] = -1; s_claimed[ev_idx[e]] = 0; removed++; }
    }
    if (removed) atomicSub(&s_nmatch, removed);
    __syncthreads();
  }
  for (int i = tid; i < n; i += kReplayThreads) {
    a.match[(long long)f * stride + i] = s_match[i];
    if (a.claimed) a.claimed[(long long)f * stride + i] = s_claimed[i];
  }
  for (int i = n + tid; i < stride; i += kReplayThreads) a.match[(long long)f * stride + i] = -1;
  if (tid == 0) a.nmatches[f] = s_nmatch;
}

// -------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(F, vpMapPoints, th), one CTA per frame.
struct SearchPointsArgs {
  const int* n_points;
  const uint8_t* in_view;
  const int* level;
  const float* view_cos;
  const float* proj_xy;
  const uint8_t* desc;
  const uint8_t* has_obs;
  int point_stride;
  float th, nn_ratio;
  int list_thresh;     // candidates above this distance can never change a decision (see host code)
  uint8_t* claimed;
  int* assign;
  int* nmatches;
  uint32_t* g_lists;   // non-null: candidate lists and counts live in HBM ([B][point_stride][kListCapPts], [B][point_stride])
  uint16_t* g_cnt;     // — used when a large local map (> ~6000 points) does not fit the CTA's shared memory
};

__global__ void __launch_bounds__(kSearchThreads) k_search_points(cmos_camera cam, const cmos_keypoint* __restrict__ kps,
                                                                 const uint8_t* __restrict__ desc,
                                                                 const int* __restrict__ counts, int stride,
                                                                 const int* __restrict__ grid_start,
                                                                 const int* __restrict__ grid_idx, int max_kp,
                                                                 SearchPointsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(counts[f], stride);
  const int np = min(a.n_points[f], a.point_stride);
  // carve: lists[np][kListCapPts] u32 | assign[n] i32 | cnt[np] u16 | claimed[n] u8
  const bool in_hbm = a.g_lists != nullptr;
  uint32_t* lists = in_hbm ? a.g_lists + (size_t)f * a.point_stride * kListCapPts : (uint32_t*)smem_raw;
  int* s_assign = in_hbm ? (int*)smem_raw : (int*)(lists + (size_t)a.point_stride * kListCapPts);
  uint16_t* cnt = in_hbm ? a.g_cnt + (size_t)f * a.point_stride : (uint16_t*)(s_assign + stride);
  uint8_t* s_claimed = in_hbm ? (uint8_t*)(s_assign + stride) : (uint8_t*)(cnt + a.point_stride);
  __shared__ int s_nmatch;

  FrameDev F{kps + (long long)f * stride, desc + (long long)f * stride * 32,
             grid_start + (long long)f * (kCells + 1), grid_idx + (long long)f * max_kp, n};
  const long long pb = (long long)f * a.point_stride;
  const bool b_factor = a.th != 1.0f;

  for (int i = tid; i < n; i += kSearchThreads) {
    s_assign[i] = -1;
    s_claimed[i] = a.claimed ? a.claimed[(long long)f * stride + i] : 0;
  }
  if (tid == 0) s_nmatch = 0;

  auto query = [&](int p, float& x, float& y, float& r, int& lvl) {
    lvl = a.level[pb + p];
    r = ((double)a.view_cos[pb + p] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos, :121-126
    if (b_factor) r *= a.th;
    r = r * cam.scale_factors[lvl];
    x = a.proj_xy[2 * (pb + p)];
    y = a.proj_xy[2 * (pb + p) + 1];
  };

  for (int p = warp; p < np; p += kSearchThreads / 32) {
    int total = 0;
    if (a.in_view[pb + p]) {
      float x, y, r; int lvl;
      query(p, x, y, r, lvl);
      const Window w = make_window(cam, x, y, r);
      if (w.ok) {
        uint32_t dq[8];
#pragma unroll
        for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(a.desc + 32 * (size_t)(pb + p)) + i);
        walk_window(F, w, x, y, r, lvl - 1, lvl, lane, [&](int idx, bool pass) {
          int d = 256;
          if (pass) d = hamming256(dq, F.desc + 32 * (size_t)idx);
          const bool keep = pass && d <= a.list_thresh;
          const unsigned m = __ballot_sync(0xffffffffu, keep);
          const int pos = total + __popc(m & ((1u << lane) - 1));
          if (keep && pos < kListCapPts)
            lists[(size_t)p * kListCapPts + pos] = ((uint32_t)d << 20) | ((uint32_t)F.kps[idx].octave << 16) | (uint32_t)idx;
          total += __popc(m);
        });
      }
    }
    if (lane == 0) cnt[p] = (uint16_t)min(total, 65535);
  }
  __syncthreads();

  if (warp == 0) {
    for (int p = 0; p < np; p++) {
      const int c = cnt[p];
      if (c == 0) continue;
      int b1 = 256, l1 = -1, b2 = 256, l2 = -1, bi = -1;
      if (c <= kListCapPts) {
        unsigned key = 0xffffffffu;
        uint32_t e = 0;
        if (lane < c) {
          e = lists[(size_t)p * kListCapPts + lane];
          if (!s_claimed[e & 0xffff]) key = ((e >> 20) << 8) | (unsigned)lane;
        }
        unsigned m1 = key;
#pragma unroll
        for (int o = 16; o; o >>= 1) m1 = min(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        if (m1 != 0xffffffffu) {
          const int src = m1 & 31;
          const uint32_t e1 = __shfl_sync(0xffffffffu, e, src);
          b1 = e1 >> 20; l1 = (e1 >> 16) & 15; bi = e1 & 0xffff;
          unsigned k2 = (lane == src) ? 0xffffffffu : key;
#pragma unroll
          for (int o = 16; o; o >>= 1) k2 = min(k2, __shfl_xor_sync(0xffffffffu, k2, o));
          if (k2 != 0xffffffffu) {
            const uint32_t e2 = __shfl_sync(0xffffffffu, e, k2 & 31);
            b2 = e2 >> 20; l2 = (e2 >> 16) & 15;
          }
        }
      } else {
        float x, y, r; int lvl;
        query(p, x, y, r, lvl);
        const Window w = make_window(cam, x, y, r);
        uint32_t dq[8];
#pragma unroll
        for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(a.desc + 32 * (size_t)(pb + p)) + i);
        // stable top-2 by (distance, candidate order): merge each 32-chunk's two minima into the running pair
        walk_window(F, w, x, y, r, lvl - 1, lvl, lane, [&](int idx, bool pass) {
          int d = 256;
          if (pass && !s_claimed[idx]) d = hamming256(dq, F.desc + 32 * (size_t)idx);
          const int oct = pass ? F.kps[idx].octave : -1;
          unsigned key = ((unsigned)d << 8) | (unsigned)lane;
          for (int round = 0; round < 2; round++) {
            unsigned mk = key;
#pragma unroll
            for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
            const int cd = (int)(mk >> 8), src = mk & 31;
            if (cd >= 256) break;
            const int ci = __shfl_sync(0xffffffffu, idx, src), co = __shfl_sync(0xffffffffu, oct, src);
            if (cd < b1) { b2 = b1; l2 = l1; b1 = cd; l1 = co; bi = ci; }
            else if (cd < b2) { b2 = cd; l2 = co; }
            if (lane == src) key = 0xffffffffu;
          }
        });
      }
      if (b1 <= CMOS_TH_HIGH) {
        if (l1 == l2 && (float)b1 > a.nn_ratio * (float)b2) continue;
        if (lane == 0) {
          s_assign[bi] = p;
          s_claimed[bi] = a.has_obs[pb + p];
          s_nmatch++;
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += kSearchThreads) {
    a.assign[(long long)f * stride + i] = s_assign[i];
    if (a.claimed) a.claimed[(long long)f * stride + i] = s_claimed[i];
  }
  for (int i = n + tid; i < stride; i += kSearchThreads) a.assign[(long long)f * stride + i] = -1;
  if (tid == 0) a.nmatches[f] = s_nmatch;
}

// -------------------------------------------------------------------------------------------------
// Frame::isInFrustum (Frame.cc:191-241) + MapPoint::PredictScale (MapPoint.cc:405-420), thread per point.
__global__ void __launch_bounds__(256) k_in_frustum(cmos_camera cam, const double* __restrict__ pose15, float cos_limit,
                                                    const int* __restrict__ n_points, const double* __restrict__ xw,
                                                    const double* __restrict__ normal, const float* __restrict__ min_d,
                                                    const float* __restrict__ max_d, int point_stride,
                                                    uint8_t* __restrict__ in_view, float* __restrict__ proj_xy,
                                                    int* __restrict__ level, float* __restrict__ view_cos) {
  const int f = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= point_stride) return;
  const long long i = (long long)f * point_stride + p;
  in_view[i] = 0;
  if (p >= n_points[f]) return;
  const double* R = pose15 + 15 * f; const double* t = R + 9; const double* Ow = R + 12;
  const double* P = xw + 3 * i;
  const double pcx = (R[0] * P[0] + R[1] * P[1]) + R[2] * P[2] + t[0];
  const double pcy = (R[3] * P[0] + R[4] * P[1]) + R[5] * P[2] + t[1];
  const double pcz = (R[6] * P[0] + R[7] * P[1]) + R[8] * P[2] + t[2];
  const float PcX = (float)pcx, PcY = (float)pcy, PcZ = (float)pcz;
  if (PcZ < 0.0f) return;
  const float invz = 1.0f / PcZ;
  const float u = cam.fx * PcX * invz + cam.cx, v = cam.fy * PcY * invz + cam.cy;
  if (u < cam.min_x || u > cam.max_x) return;
  if (v < cam.min_y || v > cam.max_y) return;
  const float maxd = 1.2f * max_d[i], mind = 0.8f * min_d[i];
  const double po0 = P[0] - Ow[0], po1 = P[1] - Ow[1], po2 = P[2] - Ow[2];
  const float dist = (float)sqrt((po0 * po0 + po1 * po1) + po2 * po2);
  if (dist < mind || dist > maxd) return;
  const double* Pn = normal + 3 * i;
  const float vc = (float)(((po0 * Pn[0] + po1 * Pn[1]) + po2 * Pn[2]) / (double)dist);
  if (vc < cos_limit) return;
  const float ratio = max_d[i] / dist;
  int ns = (int)ceilf((float)log((double)ratio) / cam.log_scale_factor);
  if (ns < 0) ns = 0;
  else if (ns >= cam.nlevels) ns = cam.nlevels - 1;
  in_view[i] = 1;
  proj_xy[2 * i] = u;
  proj_xy[2 * i + 1] = v;
  level[i] = ns;
  view_cos[i] = vc;
}

// -------------------------------------------------------------------------------------------------
// Frame::UndistortKeyPoints (Frame.cc:329-355): cv::undistortPoints(mat, mat, K, dist, Mat(), K) on the keypoint
// coordinates — five fixed-point iterations of the inverse Brown model in double (OpenCV 4.13
// cvUndistortPointsInternal, criteria MAX_ITER 5), R = I, P = K.  One thread per keypoint; the other keypoint fields
// are copied.  This translation unit is compiled without FMA contraction, as the CPU code is.
struct UndistortArgs { double fx, fy, cx, cy, k[14]; };

__device__ __forceinline__ void undistort_point(const UndistortArgs& a, float xin, float yin, float* xo, float* yo) {
  double x = xin, y = yin;
  const double u = x, v = y, ifx = 1. / a.fx, ify = 1. / a.fy;
  x = (x - a.cx) * ifx;
  y = (y - a.cy) * ify;
  const double x0 = x, y0 = y;
  const double* k = a.k;
  for (int j = 0; j < 5; j++) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) { x = (u - a.cx) * ifx; y = (v - a.cy) * ify; break; }
    const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  const double xx = a.fx * x + 0 * y + a.cx, yy = 0 * x + a.fy * y + a.cy, ww = 1. / (0 * x + 0 * y + 1);
  *xo = (float)(xx * ww);
  *yo = (float)(yy * ww);
}

__global__ void __launch_bounds__(256) k_undistort(UndistortArgs a, const cmos_keypoint* __restrict__ in,
                                                   const int* __restrict__ counts, int stride,
                                                   cmos_keypoint* __restrict__ out) {
  const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= min(counts[f], stride)) return;
  cmos_keypoint kp = in[(size_t)f * stride + i];
  undistort_point(a, kp.x, kp.y, &kp.x, &kp.y);
  out[(size_t)f * stride + i] = kp;
}

__global__ void k_undistort_xy(UndistortArgs a, const float* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) undistort_point(a, in[2 * i], in[2 * i + 1], out + 2 * i, out + 2 * i + 1);
}

}  // namespace cmos

using namespace cmos;

// =================================================================================================
// Host side
// =================================================================================================
struct cmos_match {
  cmos_match_params p{};
  cudaStream_t stream = nullptr;
  cmos_camera cam{};
  // bound current frames
  const cmos_keypoint* kps = nullptr;
  const uint8_t* desc = nullptr;
  const int* counts = nullptr;
  int n_frames = 0, stride = 0;
  bool bound = false;
  int launches = 0;
  // own device buffers
  int *d_grid_start = nullptr, *d_grid_idx = nullptr;
  float4* d_cells = nullptr;          // keypoints in cell order (FrameDev::cells)
  uint32_t* g_lists = nullptr;     // HBM candidate lists of k_search_points (allocated on first use by a large local map)
  uint16_t* g_cnt = nullptr;
  uint32_t *d_lists = nullptr, *d_best = nullptr;   // SearchByProjection(frame,last) phase-1 results
  uint16_t* d_cnt = nullptr;
  // staging for host callers
  cmos_keypoint *s_kps = nullptr, *s_last_kps = nullptr;
  uint8_t *s_desc = nullptr, *s_last_desc = nullptr, *s_last_flags = nullptr, *s_claimed = nullptr;
  int *s_counts = nullptr, *s_last_counts = nullptr, *s_match = nullptr, *s_nmatches = nullptr;
  double *s_last_xw = nullptr, *s_T = nullptr;
  // points staging
  int *s_np = nullptr, *s_level = nullptr;
  uint8_t *s_in_view = nullptr, *s_pdesc = nullptr, *s_has_obs = nullptr;
  float *s_view_cos = nullptr, *s_proj = nullptr, *s_mind = nullptr, *s_maxd = nullptr;
  double *s_pxw = nullptr, *s_pnormal = nullptr, *s_pose = nullptr;
  StageTimer timer[4];   // 0 grid, 1 search(frame,last), 2 search(frame,points), 3 isInFrustum
};

namespace {
template <typename T>
int h2d(T* dst, const T* src, size_t n, cudaStream_t st) {
  CMOS_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return CMOS_OK;
}
template <typename T>
int d2h(T* dst, const T* src, size_t n, cudaStream_t st) {
  CMOS_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
  return CMOS_OK;
}
constexpr size_t kSearchSmemLimit = 220 * 1024;
size_t search_frame_smem(int last_stride, int stride, bool stage_lists = true) {
  // lists + best | match | ev_idx, ev_q, cnt | claimed | ev_bin
  return (size_t)last_stride * ((stage_lists ? kListCap : 0) + 1) * 4 + (size_t)stride * 4 + (size_t)last_stride * 3 * 2 + stride + last_stride +
         last_stride + 4 + (size_t)stride * 4 + (size_t)last_stride * 2 + 16;   // + ptr | owner | ext
}
size_t search_points_smem(int point_stride, int stride) {
  return (size_t)point_stride * kListCapPts * 4 + (size_t)stride * 4 + (size_t)point_stride * 2 + stride + 16;
}
}  // namespace

extern "C" {

int cmos_camera_init(cmos_camera* cam, int32_t width, int32_t height, float fx, float fy, float cx, float cy,
                     const float* scale_factors, int32_t nlevels, float scale_factor) {
  CMOS_REQUIRE(cam && scale_factors && nlevels >= 1 && nlevels <= CMOS_MAX_LEVELS && width > 0 && height > 0,
               "bad argument");
  std::memset(cam, 0, sizeof(*cam));
  cam->min_x = 0.0f; cam->max_x = (float)width; cam->min_y = 0.0f; cam->max_y = (float)height;
  cam->grid_element_width_inv = static_cast<float>(CMOS_GRID_COLS) / static_cast<float>(cam->max_x - cam->min_x);
  cam->grid_element_height_inv = static_cast<float>(CMOS_GRID_ROWS) / static_cast<float>(cam->max_y - cam->min_y);
  cam->fx = fx; cam->fy = fy; cam->cx = cx; cam->cy = cy;
  cam->nlevels = nlevels;
  for (int i = 0; i < nlevels; i++) cam->scale_factors[i] = scale_factors[i];
  cam->log_scale_factor = (float)std::log((double)scale_factor);   // Frame.cc:110: log(float) assigned to float
  return CMOS_OK;
}

int cmos_match_create(const cmos_match_params* params, cmos_match_t* out) {
  CMOS_REQUIRE(params && out, "null argument");
  CMOS_REQUIRE(params->max_batch > 0 && params->max_keypoints > 0 && params->max_keypoints <= 65535 &&
               params->max_points >= 0 && params->max_points <= 65535, "bad sizes");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_REQUIRE(params->device >= 0 && params->device < ndev, "device %d out of range", params->device);
  CMOS_CUDA_OK(cudaSetDevice(params->device));
  if (search_frame_smem(params->max_keypoints, params->max_keypoints, false) > kSearchSmemLimit) {
    set_error("max_keypoints too large for the search kernels' shared memory");
    return CMOS_ERR_ARG;
  }
  cmos_match* h = new cmos_match();
  h->p = *params;
  const size_t B = params->max_batch, K = params->max_keypoints, P = std::max(params->max_points, 1);
  cudaError_t err = cudaSuccess;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) err = cudaErrorUnknown;
  h->d_grid_start = dev_alloc<int>(B * (kCells + 1), &err);
  h->d_grid_idx = dev_alloc<int>(B * K, &err);
  h->d_cells = dev_alloc<float4>(B * K, &err);
  h->d_lists = dev_alloc<uint32_t>(B * K * kListCap, &err);
  h->d_best = dev_alloc<uint32_t>(B * K, &err);
  h->d_cnt = dev_alloc<uint16_t>(B * K, &err);
  h->s_kps = dev_alloc<cmos_keypoint>(B * K, &err);
  h->s_last_kps = dev_alloc<cmos_keypoint>(B * K, &err);
  h->s_desc = dev_alloc<uint8_t>(B * K * 32, &err);
  h->s_last_desc = dev_alloc<uint8_t>(B * K * 32, &err);
  h->s_last_flags = dev_alloc<uint8_t>(B * K, &err);
  h->s_claimed = dev_alloc<uint8_t>(B * K, &err);
  h->s_counts = dev_alloc<int>(B, &err);
  h->s_last_counts = dev_alloc<int>(B, &err);
  h->s_match = dev_alloc<int>(B * K, &err);
  h->s_nmatches = dev_alloc<int>(B, &err);
  h->s_last_xw = dev_alloc<double>(B * K * 3, &err);
  h->s_T = dev_alloc<double>(B * 16, &err);
  h->s_np = dev_alloc<int>(B, &err);
  h->s_level = dev_alloc<int>(B * P, &err);
  h->s_in_view = dev_alloc<uint8_t>(B * P, &err);
  h->s_pdesc = dev_alloc<uint8_t>(B * P * 32, &err);
  h->s_has_obs = dev_alloc<uint8_t>(B * P, &err);
  h->s_view_cos = dev_alloc<float>(B * P, &err);
  h->s_proj = dev_alloc<float>(B * P * 2, &err);
  h->s_mind = dev_alloc<float>(B * P, &err);
  h->s_maxd = dev_alloc<float>(B * P, &err);
  h->s_pxw = dev_alloc<double>(B * P * 3, &err);
  h->s_pnormal = dev_alloc<double>(B * P * 3, &err);
  h->s_pose = dev_alloc<double>(B * 15, &err);
  if (err != cudaSuccess) {
    set_error("device allocation failed: %s", cudaGetErrorString(err));
    cmos_match_destroy(h);
    return CMOS_ERR_CUDA;
  }
  cudaFuncSetAttribute(k_sf_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_search_points, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_build_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  CMOS_CUDA_OK(cudaGetLastError());
  *out = h;
  return CMOS_OK;
}

int cmos_match_destroy(cmos_match_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->p.device);
  void* bufs[] = {h->d_grid_start, h->d_grid_idx, h->d_cells, h->d_lists, h->d_best, h->d_cnt, h->s_kps, h->s_last_kps, h->s_desc, h->s_last_desc, h->s_last_flags,
                  h->s_claimed, h->s_counts, h->s_last_counts, h->s_match, h->s_nmatches, h->s_last_xw, h->s_T,
                  h->s_np, h->s_level, h->s_in_view, h->s_pdesc, h->s_has_obs, h->s_view_cos, h->s_proj, h->s_mind,
                  h->s_maxd, h->s_pxw, h->s_pnormal, h->s_pose, h->g_lists, h->g_cnt};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (h->stream) cudaStreamDestroy(h->stream);
  for (StageTimer& t : h->timer) t.destroy();
  delete h;
  return CMOS_OK;
}

int cmos_match_set_frames(cmos_match_t h, const cmos_camera* cam, const cmos_keypoint* keypoints,
                          const uint8_t* descriptors, const int32_t* counts, int32_t n_frames, int32_t stride,
                          int32_t on_device, void* stream) {
  CMOS_REQUIRE(h && cam && keypoints && descriptors && counts, "null argument");
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->p.max_batch, "n_frames %d outside 1..%d", n_frames, h->p.max_batch);
  CMOS_REQUIRE(stride >= 1 && stride <= h->p.max_keypoints, "stride %d outside 1..%d", stride, h->p.max_keypoints);
  CMOS_REQUIRE(cam->nlevels >= 1 && cam->nlevels <= CMOS_MAX_LEVELS, "bad camera");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  h->cam = *cam;
  if (on_device) {
    h->kps = keypoints; h->desc = descriptors; h->counts = counts;
  } else {
    const size_t n = (size_t)n_frames * stride;
    int rc;
    if ((rc = h2d(h->s_kps, keypoints, n, st))) return rc;
    if ((rc = h2d(h->s_desc, descriptors, n * 32, st))) return rc;
    if ((rc = h2d(h->s_counts, counts, n_frames, st))) return rc;
    h->kps = h->s_kps; h->desc = h->s_desc; h->counts = h->s_counts;
  }
  h->n_frames = n_frames;
  h->stride = stride;
  const size_t smem = (kCells + 1) * sizeof(int) + (size_t)stride * sizeof(short) + 16;
  NvtxRange nvtx_grid("cmos.match.build_grid");
  h->timer[0].begin(st);
  k_build_grid<<<n_frames, 256, smem, st>>>(h->cam, h->kps, h->counts, stride, h->p.max_keypoints, h->d_grid_start,
                                            h->d_grid_idx, h->d_cells);
  h->timer[0].mark(st);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  h->bound = true;
  if (!on_device) CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_match_debug_grid(cmos_match_t h, int32_t frame, int32_t* grid_start, int32_t* grid_idx) {
  CMOS_REQUIRE(h && grid_start && grid_idx && frame >= 0 && frame < h->p.max_batch, "bad argument");
  if (!h->bound) { set_error("no frames bound"); return CMOS_ERR_STATE; }
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  CMOS_CUDA_OK(cudaDeviceSynchronize());
  CMOS_CUDA_OK(cudaMemcpy(grid_start, h->d_grid_start + (size_t)frame * (kCells + 1), (kCells + 1) * sizeof(int),
                          cudaMemcpyDeviceToHost));
  CMOS_CUDA_OK(cudaMemcpy(grid_idx, h->d_grid_idx + (size_t)frame * h->p.max_keypoints, (size_t)h->stride * sizeof(int),
                          cudaMemcpyDeviceToHost));
  return CMOS_OK;
}

int cmos_match_search_by_projection_frame(cmos_match_t h, const double* Tcw, const cmos_keypoint* last_keypoints,
                                          const int32_t* last_counts, const uint8_t* last_flags,
                                          const double* last_xw, const uint8_t* last_descriptors,
                                          int32_t last_stride, float th, int32_t check_orientation,
                                          uint8_t* claimed, int32_t* match, int32_t* nmatches, int32_t on_device,
                                          void* stream) {
  CMOS_REQUIRE(h && Tcw && last_keypoints && last_counts && last_flags && last_xw && last_descriptors && match &&
               nmatches, "null argument");
  if (!h->bound) { set_error("cmos_match_set_frames must be called first"); return CMOS_ERR_STATE; }
  CMOS_REQUIRE(last_stride >= 1 && last_stride <= h->p.max_keypoints, "last_stride %d outside 1..%d", last_stride,
               h->p.max_keypoints);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const int B = h->n_frames;
  const size_t nl = (size_t)B * last_stride, nc = (size_t)B * h->stride;
  SearchFrameArgs a{};
  a.last_stride = last_stride; a.th = th; a.check_ori = check_orientation;
  { const char* e = std::getenv("CMOS_MATCH_NO_CELLS"); a.cells = (e && e[0] == '1') ? nullptr : h->d_cells; }     // (A/B knob)
  if (on_device) {
    a.Tcw = Tcw; a.last_kps = last_keypoints; a.last_counts = last_counts; a.last_flags = last_flags;
    a.last_xw = last_xw; a.last_desc = last_descriptors; a.claimed = claimed; a.match = match; a.nmatches = nmatches;
  } else {
    int rc;
    if ((rc = h2d(h->s_T, Tcw, (size_t)B * 16, st))) return rc;
    if ((rc = h2d(h->s_last_kps, last_keypoints, nl, st))) return rc;
    if ((rc = h2d(h->s_last_counts, last_counts, B, st))) return rc;
    if ((rc = h2d(h->s_last_flags, last_flags, nl, st))) return rc;
    if ((rc = h2d(h->s_last_xw, last_xw, nl * 3, st))) return rc;
    if ((rc = h2d(h->s_last_desc, last_descriptors, nl * 32, st))) return rc;
    if (claimed && (rc = h2d(h->s_claimed, claimed, nc, st))) return rc;
    a.Tcw = h->s_T; a.last_kps = h->s_last_kps; a.last_counts = h->s_last_counts; a.last_flags = h->s_last_flags;
    a.last_xw = h->s_last_xw; a.last_desc = h->s_last_desc; a.claimed = claimed ? h->s_claimed : nullptr;
    a.match = h->s_match; a.nmatches = h->s_nmatches;
  }
  NvtxRange nvtx_frame("cmos.match.search_by_projection_frame");
  h->timer[1].begin(st);
  k_sf_lists<<<dim3((last_stride + kSfWarps - 1) / kSfWarps, B), kSfWarps * 32, 0, st>>>(
      h->cam, h->kps, h->desc, h->counts, h->stride, h->d_grid_start, h->d_grid_idx, h->p.max_keypoints, a, h->d_lists,
      h->d_best, h->d_cnt);
  const bool stage = search_frame_smem(last_stride, h->stride, true) <= kSearchSmemLimit;
  k_sf_replay<<<B, kReplayThreads, search_frame_smem(last_stride, h->stride, stage), st>>>(
      h->cam, h->kps, h->desc, h->counts, h->stride, h->d_grid_start, h->d_grid_idx, h->p.max_keypoints, a, h->d_lists,
      h->d_best, h->d_cnt, stage ? 1 : 0);
  h->timer[1].mark(st);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 2;
  if (!on_device) {
    int rc;
    if ((rc = d2h(match, h->s_match, nc, st))) return rc;
    if ((rc = d2h(nmatches, h->s_nmatches, B, st))) return rc;
    if (claimed && (rc = d2h(claimed, h->s_claimed, nc, st))) return rc;
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
  }
  return CMOS_OK;
}

int cmos_match_search_by_projection_points(cmos_match_t h, const int32_t* n_points, const uint8_t* in_view,
                                           const int32_t* level, const float* view_cos, const float* proj_xy,
                                           const uint8_t* descriptors, const uint8_t* has_obs,
                                           int32_t point_stride, float th, float nn_ratio, uint8_t* claimed,
                                           int32_t* assign, int32_t* nmatches, int32_t on_device, void* stream) {
  CMOS_REQUIRE(h && n_points && in_view && level && view_cos && proj_xy && descriptors && has_obs && assign && nmatches,
               "null argument");
  if (!h->bound) { set_error("cmos_match_set_frames must be called first"); return CMOS_ERR_STATE; }
  CMOS_REQUIRE(point_stride >= 1 && point_stride <= h->p.max_points, "point_stride %d outside 1..%d", point_stride,
               h->p.max_points);
  CMOS_REQUIRE(nn_ratio > 0.f, "nn_ratio must be positive");
  // lists in shared memory when they fit (about 6000 points next to 2000 keypoints), otherwise in HBM scratch: a large
  // local map must not make the drop-in throw (ORBmatcher::SearchByProjection(F, points) passes the whole list)
  const bool lists_in_hbm = search_points_smem(point_stride, h->stride) > 220 * 1024;
  CMOS_REQUIRE(!lists_in_hbm || (size_t)h->stride * 5 + 16 <= 220 * 1024, "stride %d exceeds the search kernel's shared memory", h->stride);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const int B = h->n_frames;
  const size_t np = (size_t)B * point_stride, nc = (size_t)B * h->stride;
  SearchPointsArgs a{};
  a.point_stride = point_stride; a.th = th; a.nn_ratio = nn_ratio;
  // A second-best candidate with nn_ratio * dist > TH_HIGH can never reject a best <= TH_HIGH
  // (ORBmatcher.cc:109-110), so such candidates need not be listed; +2 is slack for float rounding.
  a.list_thresh = std::min(255, (int)std::ceil(CMOS_TH_HIGH / nn_ratio) + 2);
  if (on_device) {
    a.n_points = n_points; a.in_view = in_view; a.level = level; a.view_cos = view_cos; a.proj_xy = proj_xy;
    a.desc = descriptors; a.has_obs = has_obs; a.claimed = claimed; a.assign = assign; a.nmatches = nmatches;
  } else {
    int rc;
    if ((rc = h2d(h->s_np, n_points, B, st))) return rc;
    if ((rc = h2d(h->s_in_view, in_view, np, st))) return rc;
    if ((rc = h2d(h->s_level, level, np, st))) return rc;
    if ((rc = h2d(h->s_view_cos, view_cos, np, st))) return rc;
    if ((rc = h2d(h->s_proj, proj_xy, np * 2, st))) return rc;
    if ((rc = h2d(h->s_pdesc, descriptors, np * 32, st))) return rc;
    if ((rc = h2d(h->s_has_obs, has_obs, np, st))) return rc;
    if (claimed && (rc = h2d(h->s_claimed, claimed, nc, st))) return rc;
    a.n_points = h->s_np; a.in_view = h->s_in_view; a.level = h->s_level; a.view_cos = h->s_view_cos;
    a.proj_xy = h->s_proj; a.desc = h->s_pdesc; a.has_obs = h->s_has_obs;
    a.claimed = claimed ? h->s_claimed : nullptr; a.assign = h->s_match; a.nmatches = h->s_nmatches;
  }
  if (lists_in_hbm) {
    const size_t need = (size_t)h->p.max_batch * std::max(h->p.max_points, 1);
    if (!h->g_lists) {
      cudaError_t err = cudaSuccess;
      h->g_lists = dev_alloc<uint32_t>(need * kListCapPts, &err);
      h->g_cnt = dev_alloc<uint16_t>(need, &err);
      CMOS_CUDA_OK(err);
    }
    a.g_lists = h->g_lists; a.g_cnt = h->g_cnt;
  }
  NvtxRange nvtx_pts("cmos.match.search_by_projection_points");
  h->timer[2].begin(st);
  k_search_points<<<B, kSearchThreads, lists_in_hbm ? (size_t)h->stride * 5 + 16 : search_points_smem(point_stride, h->stride), st>>>(
      h->cam, h->kps, h->desc, h->counts, h->stride, h->d_grid_start, h->d_grid_idx, h->p.max_keypoints, a);
  h->timer[2].mark(st);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  if (!on_device) {
    int rc;
    if ((rc = d2h(assign, h->s_match, nc, st))) return rc;
    if ((rc = d2h(nmatches, h->s_nmatches, B, st))) return rc;
    if (claimed && (rc = d2h(claimed, h->s_claimed, nc, st))) return rc;
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
  }
  return CMOS_OK;
}

int cmos_match_is_in_frustum(cmos_match_t h, const cmos_camera* cam, const double* pose15, float view_cos_limit,
                             const int32_t* n_points, const double* xw, const double* normal,
                             const float* min_distance, const float* max_distance, int32_t point_stride,
                             int32_t n_frames, uint8_t* in_view, float* proj_xy, int32_t* level, float* view_cos,
                             int32_t on_device, void* stream) {
  CMOS_REQUIRE(h && cam && pose15 && n_points && xw && normal && min_distance && max_distance && in_view && proj_xy &&
               level && view_cos, "null argument");
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->p.max_batch, "n_frames %d outside 1..%d", n_frames, h->p.max_batch);
  CMOS_REQUIRE(point_stride >= 1 && point_stride <= h->p.max_points, "point_stride %d outside 1..%d", point_stride,
               h->p.max_points);
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const size_t np = (size_t)n_frames * point_stride;
  const double *d_pose = pose15, *d_xw = xw, *d_n = normal;
  const int* d_np = n_points;
  const float *d_min = min_distance, *d_max = max_distance;
  uint8_t* o_view = in_view; float* o_proj = proj_xy; int* o_level = level; float* o_cos = view_cos;
  if (!on_device) {
    int rc;
    if ((rc = h2d(h->s_pose, pose15, (size_t)n_frames * 15, st))) return rc;
    if ((rc = h2d(h->s_np, n_points, n_frames, st))) return rc;
    if ((rc = h2d(h->s_pxw, xw, np * 3, st))) return rc;
    if ((rc = h2d(h->s_pnormal, normal, np * 3, st))) return rc;
    if ((rc = h2d(h->s_mind, min_distance, np, st))) return rc;
    if ((rc = h2d(h->s_maxd, max_distance, np, st))) return rc;
    d_pose = h->s_pose; d_np = h->s_np; d_xw = h->s_pxw; d_n = h->s_pnormal; d_min = h->s_mind; d_max = h->s_maxd;
    o_view = h->s_in_view; o_proj = h->s_proj; o_level = h->s_level; o_cos = h->s_view_cos;
  }
  h->timer[3].begin(st);
  k_in_frustum<<<dim3((point_stride + 255) / 256, n_frames), 256, 0, st>>>(*cam, d_pose, view_cos_limit, d_np, d_xw, d_n,
                                                                        d_min, d_max, point_stride, o_view, o_proj,
                                                                        o_level, o_cos);
  h->timer[3].mark(st);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 1;
  if (!on_device) {
    int rc;
    if ((rc = d2h(in_view, h->s_in_view, np, st))) return rc;
    if ((rc = d2h(proj_xy, h->s_proj, np * 2, st))) return rc;
    if ((rc = d2h(level, h->s_level, np, st))) return rc;
    if ((rc = d2h(view_cos, h->s_view_cos, np, st))) return rc;
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
  }
  return CMOS_OK;
}

int cmos_match_set_profiling(cmos_match_t h, int32_t enable) {
  CMOS_REQUIRE(h, "null handle");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  for (StageTimer& t : h->timer) { t.reset(); t.enabled = enable != 0; }
  return CMOS_OK;
}

int cmos_match_stage_times(cmos_match_t h, double* ms, int64_t* calls) {
  CMOS_REQUIRE(h && ms && calls, "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  for (int i = 0; i < 4; i++) {
    h->timer[i].fold();
    ms[i] = h->timer[i].total_ms[0];
    calls[i] = h->timer[i].calls;
  }
  return CMOS_OK;
}

int cmos_match_last_launch_count(cmos_match_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

static UndistortArgs undistort_args(const float* K4, const float* dist, int n_dist) {
  UndistortArgs a{};
  a.fx = K4[0]; a.fy = K4[1]; a.cx = K4[2]; a.cy = K4[3];
  for (int i = 0; i < n_dist && i < 14; i++) a.k[i] = dist[i];
  return a;
}

int cmos_match_undistort_keypoints(cmos_match_t h, const float* K4, const float* dist_coef, int32_t n_dist,
                                   const cmos_keypoint* keypoints, const int32_t* counts, int32_t n_frames, int32_t stride,
                                   cmos_keypoint* undistorted, int32_t on_device, void* stream) {
  CMOS_REQUIRE(h && K4 && dist_coef && keypoints && counts && undistorted, "null argument");
  CMOS_REQUIRE(n_dist >= 4 && n_dist <= 14, "n_dist %d outside 4..14", n_dist);
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->p.max_batch && stride >= 1 && stride <= h->p.max_keypoints, "bad sizes");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const size_t n = (size_t)n_frames * stride;
  const cmos_keypoint* src = keypoints;
  const int* cnt = counts;
  cmos_keypoint* dst = undistorted;
  if (!on_device) {
    int rc;
    if ((rc = h2d(h->s_last_kps, keypoints, n, st))) return rc;
    if ((rc = h2d(h->s_counts, counts, n_frames, st))) return rc;
    src = h->s_last_kps; cnt = h->s_counts; dst = h->s_kps;
  }
  h->launches = 0;
  if (dist_coef[0] == 0.0f) {     // Frame.cc:330-333: undistort_keypoints_ = keypoints_
    if (dst != src) CMOS_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(cmos_keypoint), cudaMemcpyDeviceToDevice, st));
  } else {
    k_undistort<<<dim3((stride + 255) / 256, n_frames), 256, 0, st>>>(undistort_args(K4, dist_coef, n_dist), src, cnt, stride, dst);
    CMOS_CUDA_OK(cudaGetLastError());
    h->launches = 1;
  }
  if (!on_device) {
    int rc;
    if ((rc = d2h(undistorted, h->s_kps, n, st))) return rc;
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
  }
  return CMOS_OK;
}

int cmos_camera_init_distorted(cmos_camera* cam, int32_t width, int32_t height, float fx, float fy, float cx, float cy,
                               const float* dist_coef, int32_t n_dist, const float* scale_factors, int32_t nlevels,
                               float scale_factor, int32_t device) {
  int rc = cmos_camera_init(cam, width, height, fx, fy, cx, cy, scale_factors, nlevels, scale_factor);
  if (rc) return rc;
  CMOS_REQUIRE(dist_coef && n_dist >= 4 && n_dist <= 14, "bad distortion coefficients");
  if (dist_coef[0] == 0.0f) return CMOS_OK;
  // Frame::ComputeImageBounds (Frame.cc:357-385): the four image corners through cv::undistortPoints
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_CUDA_OK(cudaSetDevice(device));
  const float corners[8] = {0.f, 0.f, (float)width, 0.f, 0.f, (float)height, (float)width, (float)height};
  float out[8];
  float* d = nullptr;
  CMOS_CUDA_OK(cudaMalloc(&d, 16 * sizeof(float)));
  const float K4[4] = {fx, fy, cx, cy};
  cudaMemcpy(d, corners, sizeof(corners), cudaMemcpyHostToDevice);
  k_undistort_xy<<<1, 32>>>(undistort_args(K4, dist_coef, n_dist), d, 4, d + 8);
  cudaError_t e = cudaMemcpy(out, d + 8, sizeof(out), cudaMemcpyDeviceToHost);
  cudaFree(d);
  CMOS_CUDA_OK(e);
  cam->min_x = std::min(out[0], out[4]);
  cam->max_x = std::max(out[2], out[6]);
  cam->min_y = std::min(out[1], out[3]);
  cam->max_y = std::max(out[5], out[7]);
  cam->grid_element_width_inv = static_cast<float>(CMOS_GRID_COLS) / static_cast<float>(cam->max_x - cam->min_x);
  cam->grid_element_height_inv = static_cast<float>(CMOS_GRID_ROWS) / static_cast<float>(cam->max_y - cam->min_y);
  return CMOS_OK;
}

}  // extern "C"
