// The remaining ORBmatcher searches for sm_100a: relocalisation and loop-closing
// projections, SearchByBoW x2, SearchForInitialization, SearchForTriangulation, Fuse x2, SearchBySim3.
//
// Behaviour follows src/ORBmatcher.cc:128-1159,1273-1384, src/KeyFrame.cc:575-626 and src/MapPoint.cc:380-420.  The
// architecture does not: every search is split into (1) a thread-per-point projection kernel that applies the
// reference's geometric gates and leaves (u, v, radius, level) per point, (2) a warp-per-point window kernel that
// walks GetFeaturesInArea in the reference's candidate order and keeps the first minimum of the Hamming distance,
// and — only where the reference's loop is order dependent (an assigned keypoint is skipped by later points) —
// (3) a single-warp replay that accepts the precomputed best while it is still free and re-walks the window
// otherwise.  The BoW searches run one warp per common vocabulary node (claims never cross nodes).  Rotation
// histograms are filtered by one small kernel shared by all searches.
//
// These are the rare-path searches (relocalisation, loop closing, initialisation, local mapping); they are built for
// exactness and low launch counts, not tuned like the per-frame path in match.cu.
// Compiled with --fmad=false (projection arithmetic must round like the unfused CPU floats/doubles).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cmos_common.h"
#include "match_device.cuh"

namespace cmos {

constexpr int kNone = 0x7fffffff;

// ---- projection ---------------------------------------------------------------------------------------------------
enum { kProjReloc = 0, kProjKeyframe = 1, kProjSim3Pair = 2 };

struct ProjArgs {
  int mode, n;
  double Ra[9], ta[3];         // kProjSim3Pair: camera a from world
  double R[9], t[3], Ow[3];    // target camera (from world, or from camera a)
  float th;
  const uint8_t* skip;
  const double* xw;
  const double* normal;
  const float* min_d;
  const float* max_d;
};

__device__ __forceinline__ void mat3_vec(const double* R, const double* p, const double* t, double* out) {
#pragma unroll
  for (int i = 0; i < 3; i++) out[i] = (R[3 * i] * p[0] + R[3 * i + 1] * p[1]) + R[3 * i + 2] * p[2] + t[i];
}

__device__ __forceinline__ int predict_scale(const cmos_camera& cam, float max_distance, float dist) {   // MapPoint.cc:390-420
  const float ratio = max_distance / dist;
  int ns = (int)ceilf((float)log((double)ratio) / cam.log_scale_factor);
  if (ns < 0) ns = 0;
  else if (ns >= cam.nlevels) ns = cam.nlevels - 1;
  return ns;
}

// out[p] = (u, v, radius, level as float bits; level < 0 = skipped)
__global__ void __launch_bounds__(128) k_kf_project(cmos_camera cam, ProjArgs a, float4* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n) return;
  float4 q = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  do {
    if (a.skip[p]) break;
    const double X[3] = {a.xw[3 * p], a.xw[3 * p + 1], a.xw[3 * p + 2]};
    double pc[3];
    float u, v, dist;
    if (a.mode == kProjReloc) {                         // ORBmatcher.cc:1297-1322
      mat3_vec(a.R, X, a.t, pc);
      const float xc = (float)pc[0], yc = (float)pc[1];
      const float invzc = (float)(1.0 / pc[2]);
      u = cam.fx * xc * invzc + cam.cx;
      v = cam.fy * yc * invzc + cam.cy;
      if (u < cam.min_x || u > cam.max_x) break;
      if (v < cam.min_y || v > cam.max_y) break;
      const double PO[3] = {X[0] - a.Ow[0], X[1] - a.Ow[1], X[2] - a.Ow[2]};
      dist = (float)sqrt((PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2]);
    } else {
      if (a.mode == kProjSim3Pair) {                    // :1005-1007
        double pa[3];
        mat3_vec(a.Ra, X, a.ta, pa);
        mat3_vec(a.R, pa, a.t, pc);
      } else {
        mat3_vec(a.R, X, a.t, pc);                      // :287-290, :749-750, :871-874
      }
      if (pc[2] < 0.0) break;
      const float invz = (float)(1 / pc[2]);
      const float x = (float)(pc[0] * invz), y = (float)(pc[1] * invz);
      u = cam.fx * x + cam.cx;
      v = cam.fy * y + cam.cy;
      if (!(u >= cam.min_x && u < cam.max_x && v >= cam.min_y && v < cam.max_y)) break;   // KeyFrame::IsInImage
      if (a.mode == kProjSim3Pair) {
        dist = (float)sqrt((pc[0] * pc[0] + pc[1] * pc[1]) + pc[2] * pc[2]);
      } else {
        const double PO[3] = {X[0] - a.Ow[0], X[1] - a.Ow[1], X[2] - a.Ow[2]};
        dist = (float)sqrt((PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2]);
        const double* Pn = a.normal + 3 * p;
        if (((PO[0] * Pn[0] + PO[1] * Pn[1]) + PO[2] * Pn[2]) < 0.5 * dist) {
          // the distance gate comes first in the reference; it is re-checked below, the order does not matter
          break;
        }
      }
    }
    const float max_d = 1.2f * a.max_d[p], min_d = 0.8f * a.min_d[p];
    if (dist < min_d || dist > max_d) break;
    const int lvl = predict_scale(cam, a.max_d[p], dist);
    q = make_float4(u, v, a.th * cam.scale_factors[lvl], __int_as_float(lvl));
  } while (false);
  out[p] = q;
}

// ---- window search -------------------------------------------------------------------------------------------------
struct BestArgs {
  int n;                       // queries
  int level_hi;                // candidate octave range = [level-1, level+level_hi]
  int chi2;                    // Fuse(KeyFrame*, points): reject e2 * inv_sigma2[octave] > 5.99
  float inv_sigma2[CMOS_MAX_LEVELS];
  const float4* q;
  const uint8_t* desc;         // query descriptors [n][32]
  const uint8_t* blocked;      // per keypoint, may be null
  int threshold;               // accept best <= threshold (kNone = keep the raw best)
  int* best_idx;
  int* best_dist;
};

// warp-wide: best (first minimum in walk order) over the window of query (u, v, radius, lvl)
template <typename Blocked>
__device__ __forceinline__ void window_best(const cmos_camera& cam, const FrameDev& F, float u, float v, float radius,
                                            int lvl, int level_hi, const uint32_t dq[8], int chi2,
                                            const float* inv_sigma2, Blocked is_blocked, int lane, int* bd_out,
                                            int* bi_out) {
  int bd = 256, bi = -1;
  const Window w = make_window(cam, u, v, radius);
  if (w.ok) {
    walk_window(F, w, u, v, radius, lvl - 1, lvl + level_hi, lane, [&](int idx, bool pass) {
      int d = 256;
      if (pass && !is_blocked(idx)) {
        bool ok = true;
        if (chi2) {
          const cmos_keypoint* kp = F.kps + idx;
          const float ex = u - kp->x, ey = v - kp->y;
          const float e2 = ex * ex + ey * ey;
          ok = !((double)(e2 * inv_sigma2[kp->octave]) > 5.99);
        }
        if (ok) d = hamming256(dq, F.desc + 32 * (size_t)idx);
      }
      unsigned key = ((unsigned)d << 8) | (unsigned)lane, mk = key;
#pragma unroll
      for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
      const int cd = (int)(mk >> 8);
      const int ci = __shfl_sync(0xffffffffu, idx, mk & 31);
      if (cd < bd) { bd = cd; bi = ci; }
    });
  }
  *bd_out = bd;
  *bi_out = bi;
}

__global__ void __launch_bounds__(256) k_kf_best(cmos_camera cam, FrameDev F, BestArgs a) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n) return;
  const float4 q = a.q[p];
  const int lvl = __float_as_int(q.w);
  int bd = 256, bi = -1;
  if (lvl >= 0) {
    uint32_t dq[8];
#pragma unroll
    for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(a.desc + 32 * (size_t)p) + i);
    const uint8_t* blocked = a.blocked;
    window_best(cam, F, q.x, q.y, q.z, lvl, a.level_hi, dq, a.chi2, a.inv_sigma2,
                [&](int idx) { return blocked && blocked[idx]; }, lane, &bd, &bi);
  }
  if (lane == 0) {
    const bool keep = bi >= 0 && (a.threshold == kNone || bd <= a.threshold);
    a.best_idx[p] = keep ? bi : -1;
    a.best_dist[p] = bd;
  }
}

// ---- rotation-histogram filter (ComputeThreeMaxima, ORBmatcher.cc:1386-1418) ---------------------------------------
// events: (output index, bin).  Events in bins outside the three maxima clear match[idx] (and flag[idx]) and decrement
// the match count; only_if_set: SearchForInitialization's "if (vnMatches12[idx1] >= 0)" (:452-455).
__device__ void rot_filter(const int2* ev, int nev, int* match, uint8_t* flag, int only_if_set, int* nmatches,
                           int* s_hist, int* s_keep) {
  const int tid = threadIdx.x, T = blockDim.x;
  for (int i = tid; i < CMOS_HISTO_LENGTH; i += T) s_hist[i] = 0;
  __syncthreads();
  for (int e = tid; e < nev; e += T) atomicAdd(&s_hist[ev[e].y], 1);
  __syncthreads();
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
  for (int e = tid; e < nev; e += T) {
    const int b = ev[e].y, idx = ev[e].x;
    if (b == s_keep[0] || b == s_keep[1] || b == s_keep[2]) continue;
    if (only_if_set && match[idx] < 0) continue;
    match[idx] = -1;
    if (flag) flag[idx] = 0;
    removed++;
  }
  if (removed) atomicSub(nmatches, removed);
}

__global__ void __launch_bounds__(256) k_rot_filter(const int2* ev, const int* nev, int* match, uint8_t* flag,
                                                    int only_if_set, int* nmatches) {
  __shared__ int s_hist[CMOS_HISTO_LENGTH], s_keep[3];
  rot_filter(ev, *nev, match, flag, only_if_set, nmatches, s_hist, s_keep);
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {
  const float factor = 1.0f / CMOS_HISTO_LENGTH;
  float rot = a1 - a2;
  if (rot < 0.0f) rot += 360.0f;
  int bin = (int)roundf(rot * factor);
  if (bin == CMOS_HISTO_LENGTH) bin = 0;
  return bin;
}

// ---- order-dependent replay (one warp) -------------------------------------------------------------------------------
// SearchByProjection(Frame, KeyFrame, ...) :1330-1361 and SearchByProjection(KeyFrame, Scw, ...) :331-357: a keypoint
// that has received a point is skipped by every later point.  taken[] starts as the caller's has_point / matched bytes.
struct GreedyArgs {
  int n_points, n_kp, threshold, level_hi;
  const float4* q;
  const uint8_t* desc;
  const int* best_idx;
  const int* best_dist;
  const float* angle;          // per point, for the rotation histogram (null = none)
  uint8_t* taken;              // [n_kp] in/out
  int* match;                  // [n_kp] out
  int2* ev;
  int* nev;
  int* nmatches;
};

__global__ void __launch_bounds__(32) k_kf_greedy(cmos_camera cam, FrameDev F, GreedyArgs a) {
  extern __shared__ uint8_t s_taken[];
  const int lane = threadIdx.x;
  for (int i = lane; i < a.n_kp; i += 32) { s_taken[i] = a.taken[i]; a.match[i] = -1; }
  __syncwarp();
  int nm = 0, nev = 0;
  for (int p = 0; p < a.n_points; p++) {
    int bi = a.best_idx[p], bd = a.best_dist[p];
    if (bi < 0 || bd > a.threshold) continue;
    if (s_taken[bi]) {
      const float4 q = a.q[p];
      uint32_t dq[8];
#pragma unroll
      for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(a.desc + 32 * (size_t)p) + i);
      window_best(cam, F, q.x, q.y, q.z, __float_as_int(q.w), a.level_hi, dq, 0, nullptr,
                  [&](int idx) { return s_taken[idx] != 0; }, lane, &bd, &bi);
      if (bi < 0 || bd > a.threshold) continue;
    }
    if (lane == 0) {
      s_taken[bi] = 1;
      a.match[bi] = p;
      if (a.angle) a.ev[nev] = make_int2(bi, rot_bin(a.angle[p], F.kps[bi].angle));
    }
    nm++;
    nev++;
    __syncwarp();
  }
  __syncwarp();
  for (int i = lane; i < a.n_kp; i += 32) a.taken[i] = s_taken[i];
  if (lane == 0) { *a.nmatches = nm; *a.nev = a.angle ? nev : 0; }
}

// ---- bag-of-words searches (one warp per node of the first feature vector) -------------------------------------------
struct FvDev { int nn; const int* node; const int* start; const int* idx; };

struct BowArgs {
  int mode;                    // 0 KF-Frame, 1 KF-KF, 2 triangulation
  FvDev A, B;
  const uint8_t* desc1;
  const uint8_t* desc2;
  const cmos_keypoint* kps1;
  const cmos_keypoint* kps2;
  const uint8_t* valid1;       // modes 0/1: pMP && !isBad; mode 2: has_point1
  const uint8_t* valid2;       // mode 1: pMP && !isBad; mode 2: has_point2
  uint8_t* taken2;
  float nn_ratio;
  int check_ori;
  // triangulation
  double F12[9];
  float ex, ey;
  float sf2[CMOS_MAX_LEVELS], sigma2_2[CMOS_MAX_LEVELS];
  int* match;
  int2* ev;
  int* nev;
  int* nmatches;
};

__global__ void __launch_bounds__(128) k_bow_match(BowArgs a) {
  const int lane = threadIdx.x & 31;
  const int na = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (na >= a.A.nn) return;
  const int node = a.A.node[na];
  int lo = 0, hi = a.B.nn;                      // lower_bound
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.B.node[mid] < node) lo = mid + 1; else hi = mid; }
  if (lo >= a.B.nn || a.B.node[lo] != node) return;
  const int b0 = a.B.start[lo], b1 = a.B.start[lo + 1];
  int accepted = 0;
  for (int k1 = a.A.start[na]; k1 < a.A.start[na + 1]; k1++) {
    const int i1 = a.A.idx[k1];
    if (a.mode == 2 ? a.valid1[i1] != 0 : a.valid1[i1] == 0) continue;
    uint32_t dq[8];
#pragma unroll
    for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(a.desc1 + 32 * (size_t)i1) + i);
    int pick = -1;
    if (a.mode != 2) {
      int best1 = 256, best2 = 256, bi = -1;
      for (int base = b0; base < b1; base += 32) {
        const int k2 = base + lane;
        int d = 256, i2 = 0;
        if (k2 < b1) {
          i2 = a.B.idx[k2];
          if (!a.taken2[i2] && (a.mode == 0 || a.valid2[i2])) d = hamming256(dq, a.desc2 + 32 * (size_t)i2);
        }
        unsigned key = ((unsigned)d << 8) | (unsigned)lane, mk = key;
#pragma unroll
        for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
        const int d1 = (int)(mk >> 8), l1 = (int)(mk & 31);
        int m2 = lane == l1 ? 256 : d;           // second smallest of the chunk (duplicates count)
#pragma unroll
        for (int o = 16; o; o >>= 1) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        const int c2 = __shfl_sync(0xffffffffu, i2, l1);
        if (d1 < best1) { best2 = min(best1, m2); best1 = d1; bi = c2; }
        else best2 = min(best2, d1);
      }
      const bool pass = a.mode == 0 ? best1 <= CMOS_TH_LOW : best1 < CMOS_TH_LOW;
      if (pass && (float)best1 < a.nn_ratio * (float)best2) pick = bi;
    } else {
      const cmos_keypoint kp1 = a.kps1[i1];
      const float la = (float)(kp1.x * a.F12[0] + kp1.y * a.F12[3] + a.F12[6]);
      const float lb = (float)(kp1.x * a.F12[1] + kp1.y * a.F12[4] + a.F12[7]);
      const float lc = (float)(kp1.x * a.F12[2] + kp1.y * a.F12[5] + a.F12[8]);
      const float den = la * la + lb * lb;
      int best = CMOS_TH_LOW, bi = -1;
      for (int base = b0; base < b1; base += 32) {
        const int k2 = base + lane;
        int d = 256, i2 = 0;
        if (k2 < b1) {
          i2 = a.B.idx[k2];
          if (!a.valid2[i2]) {
            const int dd = hamming256(dq, a.desc2 + 32 * (size_t)i2);
            if (dd <= CMOS_TH_LOW) {
              const cmos_keypoint kp2 = a.kps2[i2];
              const float distex = a.ex - kp2.x, distey = a.ey - kp2.y;
              bool ok = !(distex * distex + distey * distey < 100 * a.sf2[kp2.octave]);
              if (ok) {
                const float num = la * kp2.x + lb * kp2.y + lc;
                ok = den != 0 && (double)(num * num / den) < 3.84 * (double)a.sigma2_2[kp2.octave];
              }
              if (ok) d = dd;
            }
          }
        }
        // minimum distance, LAST occurrence among equals ("dist > bestDist" keeps going on equality, :664)
        unsigned key = ((unsigned)d << 8) | (unsigned)(31 - lane), mk = key;
#pragma unroll
        for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
        const int d1 = (int)(mk >> 8), l1 = 31 - (int)(mk & 31);
        const int c2 = __shfl_sync(0xffffffffu, i2, l1);
        if (d1 <= best) { best = d1; bi = c2; }
      }
      pick = bi;
    }
    if (pick >= 0) {
      if (lane == 0) {
        const int out = a.mode == 0 ? pick : i1;
        if (a.mode != 2) a.taken2[pick] = 1;
        a.match[out] = a.mode == 0 ? i1 : pick;
        if (a.check_ori) a.ev[atomicAdd(a.nev, 1)] = make_int2(out, rot_bin(a.kps1[i1].angle, a.kps2[pick].angle));
      }
      accepted++;
    }
    __syncwarp();
  }
  if (lane == 0 && accepted) atomicAdd(a.nmatches, accepted);
}

// ---- SearchForInitialization :363-468 (one warp; the loop is order dependent through vMatchedDistance) -------------
struct InitArgs {
  int n1;
  const cmos_keypoint* kps1;
  const uint8_t* desc1;
  float* prev;                 // [n1][2] in/out
  float window;
  float nn_ratio;
  int check_ori;
  int* matched_distance;       // [n2]
  int* matches21;              // [n2]
  int* matches12;              // [n1] out
  int2* ev;
  int* nev;
  int* nmatches;
};

__global__ void __launch_bounds__(32) k_search_init(cmos_camera cam2, FrameDev F2, InitArgs a) {
  const int lane = threadIdx.x;
  for (int i = lane; i < F2.n; i += 32) { a.matched_distance[i] = kNone; a.matches21[i] = -1; }
  for (int i = lane; i < a.n1; i += 32) a.matches12[i] = -1;
  __syncwarp();
  int nm = 0, nev = 0;
  for (int i1 = 0; i1 < a.n1; i1++) {
    const int level1 = a.kps1[i1].octave;
    if (level1 > 0) continue;
    const float x = a.prev[2 * i1], y = a.prev[2 * i1 + 1];
    const Window w = make_window(cam2, x, y, a.window);
    if (!w.ok) continue;
    uint32_t dq[8];
#pragma unroll
    for (int i = 0; i < 8; i++) dq[i] = __ldg((const uint32_t*)(a.desc1 + 32 * (size_t)i1) + i);
    int best = kNone, best2 = kNone, bi = -1;
    walk_window(F2, w, x, y, a.window, level1, level1, lane, [&](int idx, bool pass) {
      int d = kNone;
      if (pass) {
        const int dd = hamming256(dq, F2.desc + 32 * (size_t)idx);
        if (!(a.matched_distance[idx] <= dd)) d = dd;
      }
      // d <= 256 or kNone: order by (d, lane) with d clipped into 9 bits
      const unsigned dk = d == kNone ? 511u : (unsigned)d;
      unsigned key = (dk << 8) | (unsigned)lane, mk = key;
#pragma unroll
      for (int o = 16; o; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o));
      const unsigned k1 = mk >> 8;
      const int l1 = (int)(mk & 31);
      unsigned m2 = lane == l1 ? 511u : dk;
#pragma unroll
      for (int o = 16; o; o >>= 1) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      const int d1 = k1 == 511u ? kNone : (int)k1, d2 = m2 == 511u ? kNone : (int)m2;
      const int c = __shfl_sync(0xffffffffu, idx, l1);
      if (d1 < best) { best2 = min(best, d2); best = d1; bi = c; }
      else best2 = min(best2, d1);
    });
    if (best <= CMOS_TH_LOW && (float)best < (float)best2 * a.nn_ratio) {
      if (lane == 0) {
        const int old = a.matches21[bi];
        if (old >= 0) a.matches12[old] = -1;
        a.matches12[i1] = bi;
        a.matches21[bi] = i1;
        a.matched_distance[bi] = best;
        if (a.check_ori) a.ev[nev] = make_int2(i1, rot_bin(a.kps1[i1].angle, F2.kps[bi].angle));
      }
      nev++;
    }
    __syncwarp();
  }
  __syncwarp();
  // nmatches = surviving assignments (every displacement was a "--", :422-425)
  for (int i = lane; i < a.n1; i += 32) nm += a.matches12[i] >= 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
  if (lane == 0) { *a.nmatches = nm; *a.nev = a.check_ori ? nev : 0; }
}

__global__ void k_init_update_prev(int n1, const int* matches12, const cmos_keypoint* kps2, float* prev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  const int m = matches12[i];
  if (m >= 0) { prev[2 * i] = kps2[m].x; prev[2 * i + 1] = kps2[m].y; }
}

// SearchBySim3's agreement check (:1143-1156)
__global__ void k_sim3_agree(int n1, const int* m1, const int* m2, int* match12, int* nfound) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  const int idx2 = m1[i];
  const bool ok = idx2 >= 0 && m2[idx2] == i;
  match12[i] = ok ? idx2 : -1;
  if (ok) atomicAdd(nfound, 1);
}

}  // namespace cmos

using namespace cmos;

// =================================================================================================
// Host side
// =================================================================================================
struct cmos_kfmatch {
  cmos_kfmatch_params p{};
  cudaStream_t stream = nullptr;
  int launches = 0;
  struct ViewSlot {
    cmos_camera cam{};        // lookup camera (KeyFrame: truncated bounds)
    int n = 0;
    bool bound = false;
    cmos_keypoint* kps = nullptr;
    uint8_t* desc = nullptr;
    int *gs = nullptr, *gi = nullptr, *count = nullptr;
    FrameDev dev() const { return FrameDev{kps, desc, gs, gi, n}; }
  } view[2];
  int P = 0;                  // point capacity = max(max_points, max_keypoints)
  // point staging (two sets: SearchBySim3 runs both directions)
  uint8_t *d_skip[2] = {}, *d_pdesc[2] = {};
  double *d_xw[2] = {}, *d_normal[2] = {};
  float *d_mind[2] = {}, *d_maxd[2] = {}, *d_angle = nullptr;
  float4* d_q[2] = {};
  int *d_best_idx[2] = {}, *d_best_dist[2] = {};
  // per-keypoint state / outputs
  uint8_t *d_flag = nullptr, *d_valid[2] = {}, *d_taken2 = nullptr;
  int *d_match = nullptr, *d_md = nullptr, *d_m21 = nullptr;
  int2* d_ev = nullptr;
  int* d_counters = nullptr;  // [0] nev, [1] nmatches
  float* d_prev = nullptr;
  // feature vectors
  int *d_node[2] = {}, *d_start[2] = {}, *d_feat[2] = {};
};

namespace {

template <typename T>
int up(T* dst, const T* src, size_t n, cudaStream_t st) {
  if (n == 0) return CMOS_OK;
  CMOS_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return CMOS_OK;
}
template <typename T>
int down(T* dst, const T* src, size_t n, cudaStream_t st) {
  if (n == 0) return CMOS_OK;
  CMOS_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
  return CMOS_OK;
}
#define TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// Scw -> Rcw, tcw, Ow (ORBmatcher.cc:267-272, 853-858): host arithmetic, same statements as the reference
void decompose_sim3(const double* S, double* R, double* t, double* Ow) {
  const float scw = (float)std::sqrt((S[0] * S[0] + S[1] * S[1]) + S[2] * S[2]);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R[3 * i + j] = S[4 * i + j] / scw;
    t[i] = S[4 * i + 3] / scw;
  }
  for (int i = 0; i < 3; i++) Ow[i] = -((R[i] * t[0] + R[3 + i] * t[1]) + R[6 + i] * t[2]);
}

int upload_points(cmos_kfmatch* h, int set, int n, const uint8_t* skip, const double* xw, const double* normal,
                  const float* min_d, const float* max_d, const uint8_t* desc) {
  cudaStream_t st = h->stream;
  TRY(up(h->d_skip[set], skip, n, st));
  TRY(up(h->d_xw[set], xw, (size_t)n * 3, st));
  if (normal) TRY(up(h->d_normal[set], normal, (size_t)n * 3, st));
  TRY(up(h->d_mind[set], min_d, n, st));
  TRY(up(h->d_maxd[set], max_d, n, st));
  TRY(up(h->d_pdesc[set], desc, (size_t)n * 32, st));
  return CMOS_OK;
}

ProjArgs proj_args(cmos_kfmatch* h, int set, int mode, int n, float th) {
  ProjArgs a{};
  a.mode = mode; a.n = n; a.th = th;
  a.skip = h->d_skip[set]; a.xw = h->d_xw[set]; a.normal = h->d_normal[set];
  a.min_d = h->d_mind[set]; a.max_d = h->d_maxd[set];
  return a;
}

int run_project_best(cmos_kfmatch* h, int set, int slot, const ProjArgs& pa, int level_hi, int chi2,
                     const float* inv_sigma2, const uint8_t* d_blocked, int threshold) {
  const auto& V = h->view[slot];
  cudaStream_t st = h->stream;
  if (pa.n == 0) return CMOS_OK;
  k_kf_project<<<(pa.n + 127) / 128, 128, 0, st>>>(V.cam, pa, h->d_q[set]);
  BestArgs b{};
  b.n = pa.n; b.level_hi = level_hi; b.chi2 = chi2;
  if (inv_sigma2) for (int i = 0; i < V.cam.nlevels; i++) b.inv_sigma2[i] = inv_sigma2[i];
  b.q = h->d_q[set]; b.desc = h->d_pdesc[set]; b.blocked = d_blocked; b.threshold = threshold;
  b.best_idx = h->d_best_idx[set]; b.best_dist = h->d_best_dist[set];
  k_kf_best<<<(pa.n + 7) / 8, 256, 0, st>>>(V.cam, V.dev(), b);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches += 2;
  return CMOS_OK;
}

int upload_fv(cmos_kfmatch* h, int side, const cmos_feature_vector* fv, int n_kp, FvDev* out) {
  CMOS_REQUIRE(fv && fv->n_nodes >= 0 && fv->n_nodes <= h->p.max_nodes, "feature vector has %d nodes, capacity %d",
               fv ? fv->n_nodes : -1, h->p.max_nodes);
  const int total = fv->n_nodes ? fv->start[fv->n_nodes] : 0;
  CMOS_REQUIRE(total >= 0 && total <= n_kp, "feature vector lists %d features, the view has %d", total, n_kp);
  for (int i = 1; i < fv->n_nodes; i++)
    CMOS_REQUIRE(fv->node_ids[i - 1] < fv->node_ids[i], "feature vector node ids must be strictly ascending");
  cudaStream_t st = h->stream;
  TRY(up(h->d_node[side], fv->node_ids, fv->n_nodes, st));
  if (fv->n_nodes) TRY(up(h->d_start[side], fv->start, (size_t)fv->n_nodes + 1, st));
  TRY(up(h->d_feat[side], fv->features, total, st));
  *out = FvDev{fv->n_nodes, h->d_node[side], h->d_start[side], h->d_feat[side]};
  return CMOS_OK;
}

}  // namespace

extern "C" {

int cmos_kfmatch_destroy(cmos_kfmatch_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->p.device);
  std::vector<void*> bufs = {h->d_angle, h->d_flag, h->d_taken2, h->d_match, h->d_md, h->d_m21, h->d_ev, h->d_counters, h->d_prev};
  for (int s = 0; s < 2; s++) {
    auto& V = h->view[s];
    for (void* b : {(void*)V.kps, (void*)V.desc, (void*)V.gs, (void*)V.gi, (void*)V.count, (void*)h->d_skip[s],
                    (void*)h->d_pdesc[s], (void*)h->d_xw[s], (void*)h->d_normal[s], (void*)h->d_mind[s],
                    (void*)h->d_maxd[s], (void*)h->d_q[s], (void*)h->d_best_idx[s], (void*)h->d_best_dist[s],
                    (void*)h->d_valid[s], (void*)h->d_node[s], (void*)h->d_start[s], (void*)h->d_feat[s]})
      bufs.push_back(b);
  }
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CMOS_OK;
}

int cmos_kfmatch_create(const cmos_kfmatch_params* params, cmos_kfmatch_t* out) {
  CMOS_REQUIRE(params && out, "null argument");
  CMOS_REQUIRE(params->max_keypoints > 0 && params->max_keypoints <= 16384 && params->max_points >= 0 &&
               params->max_nodes >= 0, "bad sizes (max_keypoints <= 16384)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_REQUIRE(params->device >= 0 && params->device < ndev, "device %d out of range", params->device);
  CMOS_CUDA_OK(cudaSetDevice(params->device));
  cmos_kfmatch* h = new cmos_kfmatch();
  h->p = *params;
  const size_t K = params->max_keypoints, P = std::max(params->max_points, params->max_keypoints), N = params->max_nodes;
  h->P = (int)P;
  cudaError_t err = cudaSuccess;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) err = cudaErrorUnknown;
  for (int s = 0; s < 2; s++) {
    auto& V = h->view[s];
    V.kps = dev_alloc<cmos_keypoint>(K, &err);
    V.desc = dev_alloc<uint8_t>(K * 32, &err);
    V.gs = dev_alloc<int>(kCells + 1, &err);
    V.gi = dev_alloc<int>(K, &err);
    V.count = dev_alloc<int>(1, &err);
    h->d_skip[s] = dev_alloc<uint8_t>(P, &err);
    h->d_pdesc[s] = dev_alloc<uint8_t>(P * 32, &err);
    h->d_xw[s] = dev_alloc<double>(P * 3, &err);
    h->d_normal[s] = dev_alloc<double>(P * 3, &err);
    h->d_mind[s] = dev_alloc<float>(P, &err);
    h->d_maxd[s] = dev_alloc<float>(P, &err);
    h->d_q[s] = dev_alloc<float4>(P, &err);
    h->d_best_idx[s] = dev_alloc<int>(P, &err);
    h->d_best_dist[s] = dev_alloc<int>(P, &err);
    h->d_valid[s] = dev_alloc<uint8_t>(K, &err);
    h->d_node[s] = dev_alloc<int>(N, &err);
    h->d_start[s] = dev_alloc<int>(N + 1, &err);
    h->d_feat[s] = dev_alloc<int>(K, &err);
  }
  h->d_angle = dev_alloc<float>(P, &err);
  h->d_flag = dev_alloc<uint8_t>(K, &err);
  h->d_taken2 = dev_alloc<uint8_t>(K, &err);
  h->d_match = dev_alloc<int>(K, &err);
  h->d_md = dev_alloc<int>(K, &err);
  h->d_m21 = dev_alloc<int>(K, &err);
  h->d_ev = dev_alloc<int2>(std::max(P, K), &err);
  h->d_counters = dev_alloc<int>(4, &err);
  h->d_prev = dev_alloc<float>(K * 2, &err);
  if (err != cudaSuccess) {
    set_error("device allocation failed: %s", cudaGetErrorString(err));
    cmos_kfmatch_destroy(h);
    return CMOS_ERR_CUDA;
  }
  cudaFuncSetAttribute(k_kf_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  *out = h;
  return CMOS_OK;
}

int cmos_kfmatch_set_view(cmos_kfmatch_t h, int32_t slot, const cmos_camera* cam, int32_t is_keyframe,
                          const cmos_keypoint* keypoints, const uint8_t* descriptors, int32_t n) {
  CMOS_REQUIRE(h && cam && (slot == 0 || slot == 1), "bad argument");
  CMOS_REQUIRE(n >= 0 && n <= h->p.max_keypoints && (n == 0 || (keypoints && descriptors)), "n %d outside 0..%d", n,
               h->p.max_keypoints);
  CMOS_REQUIRE(cam->nlevels >= 1 && cam->nlevels <= CMOS_MAX_LEVELS, "bad camera");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  auto& V = h->view[slot];
  cudaStream_t st = h->stream;
  TRY(up(V.kps, keypoints, n, st));
  TRY(up(V.desc, descriptors, (size_t)n * 32, st));
  TRY(up(V.count, &n, 1, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));   // &n is a stack variable
  // the grid is the Frame's (float bounds, Frame.cc:158-173, copied by KeyFrame.cc:98-104)
  TRY(launch_build_grid(*cam, V.kps, V.count, std::max(n, 1), h->p.max_keypoints, V.gs, V.gi, 1, st));
  V.cam = *cam;
  if (is_keyframe) {          // KeyFrame.h:179-182: const int min_x_, min_y_, max_x_, max_y_
    V.cam.min_x = (float)(int)cam->min_x; V.cam.max_x = (float)(int)cam->max_x;
    V.cam.min_y = (float)(int)cam->min_y; V.cam.max_y = (float)(int)cam->max_y;
  }
  V.n = n;
  V.bound = true;
  h->launches = 1;
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_kfmatch_search_by_projection_reloc(cmos_kfmatch_t h, const double* Tcw, int32_t n_kf, const uint8_t* kf_valid,
                                            const double* kf_xw, const float* kf_min_distance,
                                            const float* kf_max_distance, const uint8_t* kf_descriptors,
                                            const float* kf_angle, float th, int32_t orb_dist,
                                            int32_t check_orientation, uint8_t* cur_has_point, int32_t* cur_match,
                                            int32_t* nmatches) {
  CMOS_REQUIRE(h && Tcw && cur_has_point && cur_match && nmatches, "null argument");
  CMOS_REQUIRE(h->view[0].bound, "cmos_kfmatch_set_view(slot 0) must be called first");
  CMOS_REQUIRE(n_kf >= 0 && n_kf <= h->P, "n_kf %d outside 0..%d", n_kf, h->P);
  CMOS_REQUIRE(n_kf == 0 || (kf_valid && kf_xw && kf_min_distance && kf_max_distance && kf_descriptors && kf_angle),
               "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  const auto& V = h->view[0];
  h->launches = 0;
  std::vector<uint8_t> skip(n_kf);
  for (int i = 0; i < n_kf; i++) skip[i] = !kf_valid[i];
  TRY(upload_points(h, 0, n_kf, skip.data(), kf_xw, nullptr, kf_min_distance, kf_max_distance, kf_descriptors));
  TRY(up(h->d_angle, kf_angle, n_kf, st));
  TRY(up(h->d_flag, cur_has_point, V.n, st));
  ProjArgs pa = proj_args(h, 0, kProjReloc, n_kf, th);
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) pa.R[3 * i + j] = Tcw[4 * i + j]; pa.t[i] = Tcw[4 * i + 3]; }
  for (int i = 0; i < 3; i++) pa.Ow[i] = -((pa.R[i] * pa.t[0] + pa.R[3 + i] * pa.t[1]) + pa.R[6 + i] * pa.t[2]);
  TRY(run_project_best(h, 0, 0, pa, 1, 0, nullptr, h->d_flag, kNone));
  GreedyArgs g{};
  g.n_points = n_kf; g.n_kp = V.n; g.threshold = orb_dist; g.level_hi = 1;
  g.q = h->d_q[0]; g.desc = h->d_pdesc[0]; g.best_idx = h->d_best_idx[0]; g.best_dist = h->d_best_dist[0];
  g.angle = check_orientation ? h->d_angle : nullptr;
  g.taken = h->d_flag; g.match = h->d_match; g.ev = h->d_ev; g.nev = h->d_counters; g.nmatches = h->d_counters + 1;
  k_kf_greedy<<<1, 32, (size_t)V.n + 16, st>>>(V.cam, V.dev(), g);
  k_rot_filter<<<1, 256, 0, st>>>(h->d_ev, h->d_counters, h->d_match, h->d_flag, 0, h->d_counters + 1);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches += 2;
  TRY(down(cur_has_point, h->d_flag, V.n, st));
  TRY(down(cur_match, h->d_match, V.n, st));
  TRY(down(nmatches, h->d_counters + 1, 1, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_kfmatch_search_by_projection_sim3(cmos_kfmatch_t h, const double* Scw, int32_t n_points, const uint8_t* pt_skip,
                                           const double* xw, const double* normal, const float* min_distance,
                                           const float* max_distance, const uint8_t* pt_descriptors, int32_t th,
                                           uint8_t* matched, int32_t* assign, int32_t* nmatches) {
  CMOS_REQUIRE(h && Scw && matched && assign && nmatches, "null argument");
  CMOS_REQUIRE(h->view[0].bound, "cmos_kfmatch_set_view(slot 0) must be called first");
  CMOS_REQUIRE(n_points >= 0 && n_points <= h->P, "n_points %d outside 0..%d", n_points, h->P);
  CMOS_REQUIRE(n_points == 0 || (pt_skip && xw && normal && min_distance && max_distance && pt_descriptors), "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  const auto& V = h->view[0];
  h->launches = 0;
  TRY(upload_points(h, 0, n_points, pt_skip, xw, normal, min_distance, max_distance, pt_descriptors));
  TRY(up(h->d_flag, matched, V.n, st));
  ProjArgs pa = proj_args(h, 0, kProjKeyframe, n_points, (float)th);
  decompose_sim3(Scw, pa.R, pa.t, pa.Ow);
  TRY(run_project_best(h, 0, 0, pa, 0, 0, nullptr, h->d_flag, kNone));
  GreedyArgs g{};
  g.n_points = n_points; g.n_kp = V.n; g.threshold = CMOS_TH_LOW; g.level_hi = 0;
  g.q = h->d_q[0]; g.desc = h->d_pdesc[0]; g.best_idx = h->d_best_idx[0]; g.best_dist = h->d_best_dist[0];
  g.angle = nullptr;
  g.taken = h->d_flag; g.match = h->d_match; g.ev = h->d_ev; g.nev = h->d_counters; g.nmatches = h->d_counters + 1;
  k_kf_greedy<<<1, 32, (size_t)V.n + 16, st>>>(V.cam, V.dev(), g);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches += 1;
  TRY(down(matched, h->d_flag, V.n, st));
  TRY(down(assign, h->d_match, V.n, st));
  TRY(down(nmatches, h->d_counters + 1, 1, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_kfmatch_fuse(cmos_kfmatch_t h, int32_t sim3, const double* pose, const float* inv_level_sigma2, int32_t n_points,
                      const uint8_t* pt_skip, const double* xw, const double* normal, const float* min_distance,
                      const float* max_distance, const uint8_t* pt_descriptors, float th, int32_t* best_idx,
                      int32_t* best_dist, int32_t* n_fused) {
  CMOS_REQUIRE(h && pose && best_idx && best_dist && n_fused && (sim3 || inv_level_sigma2), "null argument");
  CMOS_REQUIRE(h->view[0].bound, "cmos_kfmatch_set_view(slot 0) must be called first");
  CMOS_REQUIRE(n_points >= 0 && n_points <= h->P, "n_points %d outside 0..%d", n_points, h->P);
  CMOS_REQUIRE(n_points == 0 || (pt_skip && xw && normal && min_distance && max_distance && pt_descriptors), "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  h->launches = 0;
  TRY(upload_points(h, 0, n_points, pt_skip, xw, normal, min_distance, max_distance, pt_descriptors));
  ProjArgs pa = proj_args(h, 0, kProjKeyframe, n_points, th);
  if (sim3) decompose_sim3(pose, pa.R, pa.t, pa.Ow);
  else { std::memcpy(pa.R, pose, 72); std::memcpy(pa.t, pose + 9, 24); std::memcpy(pa.Ow, pose + 12, 24); }
  TRY(run_project_best(h, 0, 0, pa, 0, sim3 ? 0 : 1, inv_level_sigma2, nullptr, CMOS_TH_LOW));
  TRY(down(best_idx, h->d_best_idx[0], n_points, st));
  TRY(down(best_dist, h->d_best_dist[0], n_points, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  int n = 0;
  for (int i = 0; i < n_points; i++) {
    if (best_idx[i] < 0) best_dist[i] = 256;
    else n++;
  }
  *n_fused = n;
  return CMOS_OK;
}

int cmos_kfmatch_search_by_sim3(cmos_kfmatch_t h, const double* pose1, const double* pose2, float s12, const double* R12,
                                const double* t12, const uint8_t* valid1, const uint8_t* already1, const double* xw1,
                                const float* min_distance1, const float* max_distance1, const uint8_t* mp_descriptors1,
                                const uint8_t* valid2, const uint8_t* already2, const double* xw2,
                                const float* min_distance2, const float* max_distance2, const uint8_t* mp_descriptors2,
                                float th, int32_t* match12, int32_t* n_found) {
  CMOS_REQUIRE(h && pose1 && pose2 && R12 && t12 && match12 && n_found, "null argument");
  CMOS_REQUIRE(h->view[0].bound && h->view[1].bound, "both view slots must be bound first");
  const int n1 = h->view[0].n, n2 = h->view[1].n;
  CMOS_REQUIRE((n1 == 0 || (valid1 && already1 && xw1 && min_distance1 && max_distance1 && mp_descriptors1)) &&
               (n2 == 0 || (valid2 && already2 && xw2 && min_distance2 && max_distance2 && mp_descriptors2)), "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  h->launches = 0;
  std::vector<uint8_t> skip1(n1), skip2(n2);
  for (int i = 0; i < n1; i++) skip1[i] = !valid1[i] || already1[i];
  for (int i = 0; i < n2; i++) skip2[i] = !valid2[i] || already2[i];
  TRY(upload_points(h, 0, n1, skip1.data(), xw1, nullptr, min_distance1, max_distance1, mp_descriptors1));
  TRY(upload_points(h, 1, n2, skip2.data(), xw2, nullptr, min_distance2, max_distance2, mp_descriptors2));
  // transformation between the cameras, ORBmatcher.cc:973-976
  double sR12[9], sR21[9], t21[3];
  const double inv_s = 1.0 / s12;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { sR12[3 * i + j] = s12 * R12[3 * i + j]; sR21[3 * i + j] = inv_s * R12[3 * j + i]; }
  for (int i = 0; i < 3; i++) t21[i] = -((sR21[3 * i] * t12[0] + sR21[3 * i + 1] * t12[1]) + sR21[3 * i + 2] * t12[2]);
  // both directions project with pKF1's intrinsics (:960-963)
  cmos_camera cam2 = h->view[1].cam;
  const cmos_camera& cam1 = h->view[0].cam;
  cam2.fx = cam1.fx; cam2.fy = cam1.fy; cam2.cx = cam1.cx; cam2.cy = cam1.cy;
  const cmos_camera saved2 = h->view[1].cam;
  h->view[1].cam = cam2;
  ProjArgs a12 = proj_args(h, 0, kProjSim3Pair, n1, th);      // KF1's points into KF2
  std::memcpy(a12.Ra, pose1, 72); std::memcpy(a12.ta, pose1 + 9, 24);
  std::memcpy(a12.R, sR21, 72); std::memcpy(a12.t, t21, 24);
  int rc = run_project_best(h, 0, 1, a12, 0, 0, nullptr, nullptr, CMOS_TH_HIGH);
  h->view[1].cam = saved2;
  if (rc) return rc;
  ProjArgs a21 = proj_args(h, 1, kProjSim3Pair, n2, th);      // KF2's points into KF1
  std::memcpy(a21.Ra, pose2, 72); std::memcpy(a21.ta, pose2 + 9, 24);
  std::memcpy(a21.R, sR12, 72); std::memcpy(a21.t, t12, 24);
  TRY(run_project_best(h, 1, 0, a21, 0, 0, nullptr, nullptr, CMOS_TH_HIGH));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(int), st));
  if (n1 > 0) {
    k_sim3_agree<<<(n1 + 255) / 256, 256, 0, st>>>(n1, h->d_best_idx[0], h->d_best_idx[1], h->d_match, h->d_counters);
    CMOS_CUDA_OK(cudaGetLastError());
    h->launches += 1;
  }
  TRY(down(match12, h->d_match, n1, st));
  TRY(down(n_found, h->d_counters, 1, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

static int bow_common(cmos_kfmatch_t h, int mode, const uint8_t* valid1, const cmos_feature_vector* fv1,
                      const uint8_t* valid2, const cmos_feature_vector* fv2, BowArgs& a, int32_t* match, int32_t* nmatches) {
  const int n1 = h->view[0].n, n2 = h->view[1].n;
  cudaStream_t st = h->stream;
  h->launches = 0;
  TRY(upload_fv(h, 0, fv1, n1, &a.A));
  TRY(upload_fv(h, 1, fv2, n2, &a.B));
  TRY(up(h->d_valid[0], valid1, n1, st));
  if (valid2) TRY(up(h->d_valid[1], valid2, n2, st));
  const int n_out = mode == 0 ? n2 : n1;
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_taken2, 0, (size_t)std::max(n2, 1), st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_match, 0xff, (size_t)std::max(n_out, 1) * sizeof(int), st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(int), st));
  a.mode = mode;
  a.desc1 = h->view[0].desc; a.desc2 = h->view[1].desc; a.kps1 = h->view[0].kps; a.kps2 = h->view[1].kps;
  a.valid1 = h->d_valid[0]; a.valid2 = h->d_valid[1]; a.taken2 = h->d_taken2;
  a.match = h->d_match; a.ev = h->d_ev; a.nev = h->d_counters; a.nmatches = h->d_counters + 1;
  if (a.A.nn > 0 && a.B.nn > 0) {
    k_bow_match<<<(a.A.nn + 3) / 4, 128, 0, st>>>(a);
    h->launches += 1;
    if (a.check_ori) {
      k_rot_filter<<<1, 256, 0, st>>>(h->d_ev, h->d_counters, h->d_match, nullptr, 0, h->d_counters + 1);
      h->launches += 1;
    }
    CMOS_CUDA_OK(cudaGetLastError());
  }
  TRY(down(match, h->d_match, n_out, st));
  TRY(down(nmatches, h->d_counters + 1, 1, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_kfmatch_search_by_bow(cmos_kfmatch_t h, int32_t mode, const uint8_t* valid1, const cmos_feature_vector* fv1,
                               const uint8_t* valid2, const cmos_feature_vector* fv2, float nn_ratio,
                               int32_t check_orientation, int32_t* match, int32_t* nmatches) {
  CMOS_REQUIRE(h && (mode == 0 || mode == 1) && fv1 && fv2 && match && nmatches, "bad argument");
  CMOS_REQUIRE(h->view[0].bound && h->view[1].bound, "both view slots must be bound first");
  CMOS_REQUIRE((h->view[0].n == 0 || valid1) && (mode == 0 || h->view[1].n == 0 || valid2), "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  BowArgs a{};
  a.nn_ratio = nn_ratio; a.check_ori = check_orientation;
  return bow_common(h, mode, valid1, fv1, mode == 1 ? valid2 : nullptr, fv2, a, match, nmatches);
}

int cmos_kfmatch_search_for_triangulation(cmos_kfmatch_t h, const uint8_t* has_point1, const cmos_feature_vector* fv1,
                                          const uint8_t* has_point2, const cmos_feature_vector* fv2, const double* F12,
                                          const double* Cw, const double* R2w, const double* t2w,
                                          const float* level_sigma2_2, int32_t check_orientation, int32_t* match12,
                                          int32_t* nmatches) {
  CMOS_REQUIRE(h && fv1 && fv2 && F12 && Cw && R2w && t2w && level_sigma2_2 && match12 && nmatches, "null argument");
  CMOS_REQUIRE(h->view[0].bound && h->view[1].bound, "both view slots must be bound first");
  CMOS_REQUIRE((h->view[0].n == 0 || has_point1) && (h->view[1].n == 0 || has_point2), "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  BowArgs a{};
  a.check_ori = check_orientation;
  std::memcpy(a.F12, F12, 72);
  const cmos_camera& cam2 = h->view[1].cam;
  // epipole of camera 1 in image 2, ORBmatcher.cc:588-594
  double C2[3];
  for (int i = 0; i < 3; i++) C2[i] = (R2w[3 * i] * Cw[0] + R2w[3 * i + 1] * Cw[1]) + R2w[3 * i + 2] * Cw[2] + t2w[i];
  const float invz = 1.0f / (float)C2[2];
  a.ex = (float)(cam2.fx * C2[0] * invz + cam2.cx);
  a.ey = (float)(cam2.fy * C2[1] * invz + cam2.cy);
  for (int i = 0; i < cam2.nlevels; i++) { a.sf2[i] = cam2.scale_factors[i]; a.sigma2_2[i] = level_sigma2_2[i]; }
  return bow_common(h, 2, has_point1, fv1, has_point2, fv2, a, match12, nmatches);
}

int cmos_kfmatch_search_for_initialization(cmos_kfmatch_t h, float* prev_matched, int32_t window_size, float nn_ratio,
                                           int32_t check_orientation, int32_t* matches12, int32_t* nmatches) {
  CMOS_REQUIRE(h && matches12 && nmatches, "null argument");
  CMOS_REQUIRE(h->view[0].bound && h->view[1].bound, "both view slots must be bound first");
  const int n1 = h->view[0].n;
  CMOS_REQUIRE(n1 == 0 || prev_matched, "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->p.device));
  cudaStream_t st = h->stream;
  h->launches = 0;
  TRY(up(h->d_prev, prev_matched, (size_t)n1 * 2, st));
  InitArgs a{};
  a.n1 = n1; a.kps1 = h->view[0].kps; a.desc1 = h->view[0].desc; a.prev = h->d_prev; a.window = (float)window_size;
  a.nn_ratio = nn_ratio; a.check_ori = check_orientation;
  a.matched_distance = h->d_md; a.matches21 = h->d_m21; a.matches12 = h->d_match;
  a.ev = h->d_ev; a.nev = h->d_counters; a.nmatches = h->d_counters + 1;
  k_search_init<<<1, 32, 0, st>>>(h->view[1].cam, h->view[1].dev(), a);
  k_rot_filter<<<1, 256, 0, st>>>(h->d_ev, h->d_counters, h->d_match, nullptr, 1, h->d_counters + 1);
  if (n1 > 0) k_init_update_prev<<<(n1 + 255) / 256, 256, 0, st>>>(n1, h->d_match, h->view[1].kps, h->d_prev);
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = 3;
  TRY(down(matches12, h->d_match, n1, st));
  TRY(down(prev_matched, h->d_prev, (size_t)n1 * 2, st));
  TRY(down(nmatches, h->d_counters + 1, 1, st));
  CMOS_CUDA_OK(cudaStreamSynchronize(st));
  return CMOS_OK;
}

int cmos_kfmatch_last_launch_count(cmos_kfmatch_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

}  // extern "C"
