// ORB front-end for sm_100a: pyramid -> per-cell FAST-9 + NMS -> quadtree distribution ->
// intensity-centroid angle -> 7x7 Gaussian -> steered BRIEF-256, batched over frames.
//
// Behaviour follows the reference's ORBextractor (src/ORBextractor.cc:410-470,765-853,1043-1132) and the
// OpenCV 4.13 fixed-point primitives it calls; the architecture does not: every
// stage is one batched launch over (work item, frame), the FAST threshold fallback and the quadtree run
// on the device, and nothing returns to the host between stages.
//
// Compiled with --fmad=false: the descriptor rotation and fastAtan2 must round exactly like the unfused
// float32 CPU arithmetic (SURVEY.md §7 hard part 2).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cmos_orb_pattern.h"
#include "cmos_common.h"
#include "orb_device.cuh"

namespace cmos {

// =================================================================================================
// Kernels
// =================================================================================================

__device__ __forceinline__ int reflect101(int i, int n) {
  // valid for -n < i < 2n-1 (border 19 <= level size is enforced on the host)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Level 0: input image -> bordered plane (cv::copyMakeBorder REFLECT_101, ORBextractor.cc:1127).
// One thread writes one aligned 16-byte vector of the plane.  Interior vectors (93 % of them): the sixteen source bytes sit at
// an arbitrary alignment of an arbitrary-pitch input row — five aligned word loads and four funnel shifts.  Vectors that
// touch the reflected border go byte by byte; vectors left / right of the border are zero padding.
__global__ void __launch_bounds__(256) k_level0(OrbGeom g, const uint8_t* __restrict__ images,
                                                long long frame_stride, int in_pitch,
                                                uint8_t* __restrict__ pyr) {
  const LevelGeom& L = g.lv[0];
  const int c16 = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  const int f = blockIdx.z;
  if (c16 * 16 >= L.pitch || row >= L.rows) return;
  const int sy = reflect101(row - kBorder, L.h);
  const uint8_t* srow = images + (long long)f * frame_stride + (long long)sy * in_pitch;
  const int bx0 = c16 * 16 - kXOff;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (bx0 >= 4 && bx0 + 19 < L.w) {
    // 4 <= bx0 and bx0 + 19 < w keep all five aligned words inside the row
    const uintptr_t a = (uintptr_t)(srow + bx0);
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const int sh = 8 * (int)(a & 3);
    const uint32_t v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3), v4 = __ldg(q + 4);
    w[0] = __funnelshift_r(v0, v1, sh); w[1] = __funnelshift_r(v1, v2, sh);
    w[2] = __funnelshift_r(v2, v3, sh); w[3] = __funnelshift_r(v3, v4, sh);
  } else if (bx0 + 15 >= -kBorder && bx0 < L.w + kBorder) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int bx = bx0 + 4 * k + b;
        if (bx >= -kBorder && bx < L.w + kBorder) word |= (uint32_t)__ldg(srow + reflect101(bx, L.w)) << (8 * b);
      }
      w[k] = word;
    }
  }
  *(uint4*)(pyr + (long long)f * g.frame_bytes + L.plane_off + (long long)row * L.pitch + c16 * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Level l from level l-1: cv::resize INTER_LINEAR fixed-point (Appendix A.1) fused with the
// reflect-101 border (ORBextractor.cc:1120-1123).  xtab entries: {coeff0, coeff1, src x, 0}; ytab: {src y, coeff0,
// coeff1, 0}.  One thread writes one aligned 4-byte word.  Interior words take the fast path: the four source
// pairs lie within 12 bytes, so each source row is three aligned word loads, every pair is cut out with a funnel
// shift and the horizontal pass is one DP2A (coeff pair x byte pair).
__device__ __forceinline__ uint32_t resize_vpass(int t0, int t1, int b0, int b1) {
  int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
  return (uint32_t)min(max(v, 0), 255);
}

// kRows output rows per thread: the column set-up (tables, window selectors) is paid once per thread
template <int kResizeRows>
__global__ void __launch_bounds__(256) k_resize(OrbGeom g, int level, const short4* __restrict__ xtab,
                                                const short4* __restrict__ ytab, uint8_t* __restrict__ pyr) {
  const LevelGeom& L = g.lv[level];
  const LevelGeom& S = g.lv[level - 1];
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int row_first = (blockIdx.y * blockDim.y + threadIdx.y) * kResizeRows;
  const int f = blockIdx.z;
  if (c4 * 4 >= L.pitch || row_first >= L.rows) return;
  const int row_end = min(row_first + kResizeRows, L.rows);
  uint8_t* frame = pyr + (long long)f * g.frame_bytes;
  const uint8_t* src = frame + S.plane_off + (long long)kBorder * S.pitch + kXOff;
  uint8_t* dst = frame + L.plane_off + c4 * 4;
  const int bx0 = c4 * 4 - kXOff;
  // fast path: interior word whose four source pairs start within bytes 0..7 of three aligned words (always the case up to
  // scale 4/3; a wider group — scale 1.5, 2.0 — takes the generic path below)
  bool fast = bx0 >= 0 && bx0 + 3 < L.w;
  unsigned coef[4] = {0u, 0u, 0u, 0u}, sh[4] = {0u, 0u, 0u, 0u};
  bool hi[4] = {false, false, false, false};
  int wb = 0;
  if (fast) {
    const int4 ta = __ldg((const int4*)(xtab + bx0)), tb = __ldg((const int4*)(xtab + bx0) + 1);
    coef[0] = (unsigned)ta.x; coef[1] = (unsigned)ta.z; coef[2] = (unsigned)tb.x; coef[3] = (unsigned)tb.z;
    const int sx[4] = {ta.y & 0xffff, ta.w & 0xffff, tb.y & 0xffff, tb.w & 0xffff};
    wb = sx[0] >> 2;
    if (sx[3] - 4 * wb > 7) fast = false;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int o = sx[b] - 4 * wb;           // 0..7
      hi[b] = o >= 4;
      sh[b] = 8u * (unsigned)(o & 3);
    }
  }
  if (fast) {
#pragma unroll
    for (int rr = 0; rr < kResizeRows; rr++) {
      const int row = row_first + rr;
      if (row >= row_end) break;
      const short4 ty = __ldg(ytab + reflect101(row - kBorder, L.h));
      const int sy0 = min(max((int)ty.x, 0), S.h - 1), sy1 = min(max((int)ty.x + 1, 0), S.h - 1);
      const uint32_t* p0 = (const uint32_t*)(src + (long long)sy0 * S.pitch) + wb;
      const uint32_t* p1 = (const uint32_t*)(src + (long long)sy1 * S.pitch) + wb;
      const uint32_t u0 = p0[0], u1 = p0[1], u2 = p0[2], v0 = p1[0], v1 = p1[1], v2 = p1[2];
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t w0 = __funnelshift_r(hi[b] ? u1 : u0, hi[b] ? u2 : u1, sh[b]);
        const uint32_t w1 = __funnelshift_r(hi[b] ? v1 : v0, hi[b] ? v2 : v1, sh[b]);
        const int t0 = (int)__dp2a_lo(coef[b], w0, 0u), t1 = (int)__dp2a_lo(coef[b], w1, 0u);
        word |= resize_vpass(t0, t1, ty.y, ty.z) << (8 * b);
      }
      *(uint32_t*)(dst + (long long)row * L.pitch) = word;
    }
  } else {
    for (int row = row_first; row < row_end; row++) {
      const short4 ty = __ldg(ytab + reflect101(row - kBorder, L.h));
      const int sy0 = min(max((int)ty.x, 0), S.h - 1), sy1 = min(max((int)ty.x + 1, 0), S.h - 1);
      const uint8_t* r0 = src + (long long)sy0 * S.pitch;
      const uint8_t* r1 = src + (long long)sy1 * S.pitch;
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        int bx = bx0 + b;
        if (bx >= -kBorder && bx < L.w + kBorder) {
          int dx = reflect101(bx, L.w);
          short4 tx = __ldg(xtab + dx);
          int sx0 = tx.z, sx1 = min(sx0 + 1, S.w - 1);
          int t0 = r0[sx0] * (int)tx.x + r0[sx1] * (int)tx.y;
          int t1 = r1[sx0] * (int)tx.x + r1[sx1] * (int)tx.y;
          word |= resize_vpass(t0, t1, ty.y, ty.z) << (8 * b);
        }
      }
      *(uint32_t*)(dst + (long long)row * L.pitch) = word;
    }
  }
}

// Second-generation resize: the INTERIOR of level l from the interior of level l-1, one CTA per 128 x 16 output tile.
// Round 1's k_resize spent ~65 thread-instructions per output pixel: every output row redid the horizontal pass of both of
// its source rows and the vertical pass was two 32-bit multiplies, shifts and clamps per pixel.  Here
//   * the horizontal pass runs ONCE per source row of the tile (about 1.2 source rows per output row) into shared memory as
//     u16 (t >> 4 <= 32640 fits), four columns per thread: three aligned word loads, funnel shifts, one DP2A per pixel;
//   * the vertical pass is two IMAD.HI per pixel: ((b * tt) >> 16) == umulhi(b << 16, tt), the "+ 2" rides as the addend;
//     the result cannot exceed 255 (the weights sum to 2048), so OpenCV's saturation is a no-op and is dropped;
//   * the reflect-101 borders of all levels are filled afterwards by ONE launch of k_borders (the resize itself clamps source
//     coordinates, ORBextractor.cc:1120 / OpenCV's xofs clamp, and never reads a border).
// Bit-exact with cv::resize INTER_LINEAR (tests compare every plane byte with the oracle and with the reference build).
// Tile height: the column set-up (table loads, funnel-shift amounts: ~110 instructions per thread) is paid once per tile, so a
// 128 x 16 tile (two output rows per thread) spent more on it than on the pixels: 50 thread-instructions per output pixel.
// 128 x 64 tiles (eight output rows per thread, 80 source rows = 20 KB of shared memory) amortise it four times better.
constexpr int kRzTW = 128;
__host__ __device__ constexpr int rz_max_src_rows(int th) { return th == 64 ? 80 : th == 32 ? 42 : 24; }
template <int kRzTH>
__global__ void __launch_bounds__(256) k_resize2(OrbGeom g, int level, const short4* __restrict__ xtab,
                                                 const short4* __restrict__ ytab, uint8_t* __restrict__ pyr) {
  constexpr int kRzMaxSrcRows = rz_max_src_rows(kRzTH);
  __shared__ __align__(8) uint16_t s_h[kRzMaxSrcRows][kRzTW];
  const LevelGeom& L = g.lv[level];
  const LevelGeom& S = g.lv[level - 1];
  const int tid = threadIdx.x, cg = tid & 31, rsub = tid >> 5;      // column group (4 output columns), row phase 0..7
  const int x_first = blockIdx.x * kRzTW + 4 * cg;
  const int r_first = blockIdx.y * kRzTH, r_last = min(r_first + kRzTH, L.h) - 1;
  uint8_t* frame = pyr + (long long)blockIdx.z * g.frame_bytes;
  const uint8_t* src = frame + S.plane_off + (long long)kBorder * S.pitch + kXOff;
  const int y_lo = min(max((int)__ldg(ytab + r_first).x, 0), S.h - 1);
  const int y_hi = min(max((int)__ldg(ytab + r_last).x + 1, 0), S.h - 1);
  if (x_first < L.w) {
    // column set-up: coefficient pairs and the byte offsets of the four source pairs inside three aligned words
    unsigned coef[4], sh[4];
    bool hi[4];
    int wb;
    {
      int sx[4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const short4 t = __ldg(xtab + min(x_first + b, L.w - 1));
        coef[b] = ((unsigned)(unsigned short)t.x) | ((unsigned)(unsigned short)t.y << 16);
        sx[b] = t.z;
      }
      wb = sx[0] >> 2;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int o = min(sx[b] - 4 * wb, 8);      // 0..8 (a clamped column of a partial group never exceeds it)
        hi[b] = o >= 4;
        sh[b] = 8u * (unsigned)(o & 3);
      }
    }
    for (int sr = y_lo + rsub; sr <= y_hi; sr += 8) {
      const uint32_t* p = (const uint32_t*)(src + (long long)sr * S.pitch) + wb;
      const uint32_t u0 = p[0], u1 = p[1], u2 = p[2];
      unsigned tt[4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t w = __funnelshift_r(hi[b] ? u1 : u0, hi[b] ? u2 : u1, sh[b]);
        tt[b] = __dp2a_lo(coef[b], w, 0u) >> 4;
      }
      *(uint2*)&s_h[sr - y_lo][4 * cg] = make_uint2(tt[0] | (tt[1] << 16), tt[2] | (tt[3] << 16));
    }
  }
  __syncthreads();
  if (x_first >= L.w) return;
  for (int r = r_first + rsub; r <= r_last; r += 8) {
    const short4 ty = __ldg(ytab + r);
    const int l0 = min(max((int)ty.x, 0), S.h - 1) - y_lo, l1 = min(max((int)ty.x + 1, 0), S.h - 1) - y_lo;
    const unsigned B0 = (unsigned)ty.y << 16, B1 = (unsigned)ty.z << 16;
    const uint2 h0 = *(const uint2*)&s_h[l0][4 * cg], h1 = *(const uint2*)&s_h[l1][4 * cg];
    const unsigned v0 = (__umulhi(B0, h0.x & 0xffffu) + 2u + __umulhi(B1, h1.x & 0xffffu)) >> 2;
    const unsigned v1 = (__umulhi(B0, h0.x >> 16) + 2u + __umulhi(B1, h1.x >> 16)) >> 2;
    const unsigned v2 = (__umulhi(B0, h0.y & 0xffffu) + 2u + __umulhi(B1, h1.y & 0xffffu)) >> 2;
    const unsigned v3 = (__umulhi(B0, h0.y >> 16) + 2u + __umulhi(B1, h1.y >> 16)) >> 2;
    *(uint32_t*)(frame + L.plane_off + (long long)(r + kBorder) * L.pitch + kXOff + x_first) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
  }
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101 + BORDER_ISOLATED) of levels first_level.. in ONE launch (ORBextractor.cc:1122):
// grid (work blocks, level - first_level, frame).  Work block b < n_side: the side borders of 21 interior rows (threads =
// rows x 12 words: words that hold bx in [-19, -1] or [w, w + 18]; words shared with the interior are read-modified-written);
// the remaining blocks: one of the 38 border rows each, word per thread.
constexpr int kSideRows = 21;     // 21 rows x 12 words = 252 threads of a 256-thread CTA
constexpr int kBorderRowsPerCta = 4;
__global__ void __launch_bounds__(256) k_borders(OrbGeom g, int first_level, int n_side_max, uint8_t* __restrict__ pyr) {
  const LevelGeom& L = g.lv[first_level + blockIdx.y];
  uint8_t* plane = pyr + (long long)blockIdx.z * g.frame_bytes + L.plane_off;
  const uint8_t* in = plane + (long long)kBorder * L.pitch + kXOff;      // interior origin
  int b = blockIdx.x;
  if (b < n_side_max) {
    const int row = b * kSideRows + threadIdx.x / 12, k = threadIdx.x % 12;
    if (threadIdx.x >= kSideRows * 12 || row >= L.h) return;
    // left words cover bx in [-20, -1] (5 words), right words start at the word holding bx = w
    const int w_first_right = (kXOff + L.w) >> 2;
    const int word = k < 5 ? (kXOff - 20) / 4 + k : w_first_right + (k - 5);
    if (4 * word >= L.pitch) return;
    uint32_t* dst = (uint32_t*)(plane + (long long)(row + kBorder) * L.pitch) + word;
    uint32_t v = *dst;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int bx = 4 * word + q - kXOff;
      if ((bx < 0 && bx >= -kBorder) || (bx >= L.w && bx < L.w + kBorder)) {
        const uint32_t px = in[(long long)row * L.pitch + reflect101(bx, L.w)];
        v = (v & ~(0xffu << (8 * q))) | (px << (8 * q));
      }
    }
    *dst = v;
    return;
  }
  b -= n_side_max;
  // kBorderRowsPerCta border rows per CTA (one row per group of 64 threads): 17 k one-row CTAs were launch-bound
  const int br = b * kBorderRowsPerCta + (threadIdx.x >> 6);
  if (br >= 2 * kBorder) return;
  const int prow = br < kBorder ? br : L.h + br;                    // plane row: 0..18 and h + 19 .. h + 37
  const int sy = reflect101(prow - kBorder, L.h);
  for (int word = threadIdx.x & 63; 4 * word < L.pitch; word += 64) {
    const int bx0 = 4 * word - kXOff;
    uint32_t v = 0;
    if (bx0 >= 0 && bx0 + 3 < L.w) {        // above / below the interior: a word-aligned copy of interior row sy
      v = *(const uint32_t*)(in + (long long)sy * L.pitch + bx0);
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int bx = bx0 + q;
        if (bx >= -kBorder && bx < L.w + kBorder) v |= (uint32_t)in[(long long)sy * L.pitch + reflect101(bx, L.w)] << (8 * q);
      }
    }
    *((uint32_t*)(plane + (long long)prow * L.pitch) + word) = v;
  }
}

// -------------------------------------------------------------------------------------------------
// FAST-9/16 (cv::FAST's cornerScore, Appendix A.3): m = max over the 16 arcs of 9 contiguous ring pixels of
// min(ring - c) and of min(c - ring); corner iff m > t, score m - 1.
//
// Both polarities travel in one register as two unsigned 16-bit halves, low = 256 + (ring - c),
// high = 256 - (ring - c), built by one multiply-add per ring pixel:  K - 65535*ring with
// K = 0x01000100 + 65535*c  (the integer value (256-x)*65536 + (256+x) has exactly those base-65536 digits since
// both are in [1,511]).  Arc minima are then packed VIMNMX.U16x2 / VIMNMX3.U16x2 chains — 56 min/max
// instructions per pixel instead of ~160 scalar ones, and no negation (an earlier scalar version tripped a ptxas
// VIMNMX3 negation fold).
__device__ __forceinline__ unsigned fast_pack(const uint8_t* __restrict__ p, unsigned K) {
  return K - 65535u * (unsigned)p[0];
}

// Cheap necessary condition: a 9-arc contains two adjacent compass points (0,4,8,12), so both polarities need
// (d0 | d8) & (d4 | d12) above t.  Returns true when the pixel may have m > t.
__device__ __forceinline__ bool fast_compass(const uint8_t* __restrict__ p, int tw, int t) {
  const unsigned K = 0x01000100u + 65535u * (unsigned)p[0];
  const unsigned a = __vmaxu2(fast_pack(p + 3 * tw, K), fast_pack(p - 3 * tw, K));
  const unsigned b = __vmaxu2(fast_pack(p + 3, K), fast_pack(p - 3, K));
  const unsigned m = __vminu2(a, b);
  return max(m & 0xffffu, m >> 16) > (unsigned)(256 + t);
}

__device__ __forceinline__ int fast_arc_max(const uint8_t* __restrict__ p, int tw) {
  const unsigned K = 0x01000100u + 65535u * (unsigned)p[0];
  unsigned v[16];
  v[0] = fast_pack(p + 3 * tw, K);
  v[1] = fast_pack(p + 3 * tw + 1, K);
  v[2] = fast_pack(p + 2 * tw + 2, K);
  v[3] = fast_pack(p + tw + 3, K);
  v[4] = fast_pack(p + 3, K);
  v[5] = fast_pack(p - tw + 3, K);
  v[6] = fast_pack(p - 2 * tw + 2, K);
  v[7] = fast_pack(p - 3 * tw + 1, K);
  v[8] = fast_pack(p - 3 * tw, K);
  v[9] = fast_pack(p - 3 * tw - 1, K);
  v[10] = fast_pack(p - 2 * tw - 2, K);
  v[11] = fast_pack(p - tw - 3, K);
  v[12] = fast_pack(p - 3, K);
  v[13] = fast_pack(p + tw - 3, K);
  v[14] = fast_pack(p + 2 * tw - 2, K);
  v[15] = fast_pack(p + 3 * tw - 1, K);
  unsigned lo2[16], lo4[16];
#pragma unroll
  for (int k = 0; k < 16; k++) lo2[k] = __vminu2(v[k], v[(k + 1) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) lo4[k] = __vminu2(lo2[k], lo2[(k + 2) & 15]);
  unsigned best = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    unsigned a = __vimin3_u16x2(lo4[k], lo4[(k + 4) & 15], v[(k + 8) & 15]);
    unsigned b = __vimin3_u16x2(lo4[k + 1], lo4[(k + 5) & 15], v[(k + 9) & 15]);
    best = __vimax3_u16x2(best, a, b);
  }
  return max((int)max(best & 0xffffu, best >> 16) - 256, 0);
}

// One CTA per (cell, frame): stage the (w_cell+6) x (h_cell+6) tile in shared memory; then per threshold
// (iniThFAST, and minThFAST only when the cell came out empty — ORBextractor.cc:808-816):
//   A  compass test over the detection area, survivors compacted into a pixel list (warp ballot);
//   B  exact packed arc score over the list (dense: no lane idles on a rejected pixel), scores > t into the map;
//   C  3x3 strict NMS over the list with zero outside the cell, survivors appended to the (frame, level) list.
// kTW x kTH: tile pitch / rows in shared memory, kCap: candidates kept per cell.  The 48 x 48 variant serves grids whose
// cells are at most 44 x 48 (KITTI / TUM: 37 x 46): 9 KB of shared memory per CTA instead of 25 KB, so 16 CTAs (all 64
// warps) are resident per SM instead of 9 — the kernel stalls on barriers and on the tile's global loads, which more
// resident warps cover.
#ifndef CMOS_FAST_MINBLOCKS
#define CMOS_FAST_MINBLOCKS 12     // resident 128-thread CTAs per SM asked of the compiler: 40 registers, 0.273 -> 0.262 ms per 64 frames (16: 32 registers with spills, 0.259)
#endif
// ((1 << 20) + d - 1) / d for d = 0..32: i / d == (i * kMagic20[d]) >> 20 for i < 2^13 (an integer division costs ~20
// instructions per thread; k_fast needed two per CTA pass)
__device__ __constant__ unsigned kMagic20[33] = {0u, 1048576u, 524288u, 349526u, 262144u, 209716u, 174763u, 149797u, 131072u, 116509u, 104858u, 95326u, 87382u, 80660u, 74899u, 69906u, 65536u, 61681u, 58255u, 55189u, 52429u, 49933u, 47663u, 45591u, 43691u, 41944u, 40330u, 38837u, 37450u, 36158u, 34953u, 33826u, 32768u};

template <int kThreads, int kTW, int kTH, int kCap>
#if CMOS_FAST_MINBLOCKS > 0
__global__ void __launch_bounds__(kThreads, CMOS_FAST_MINBLOCKS * 128 / kThreads) k_fast(
#else
__global__ void __launch_bounds__(kThreads) k_fast(
#endif
    OrbGeom g, const int4* __restrict__ cells,
                                                  const uint8_t* __restrict__ pyr, uint32_t* __restrict__ cand,
                                                  int* __restrict__ cand_count, int* __restrict__ overflow,
                                                  uint8_t* __restrict__ dbg, int dbg_cell) {
  __shared__ __align__(16) uint8_t tile[kTH * kTW];
  __shared__ __align__(16) uint8_t score[kTH * kTW];
  __shared__ uint16_t s_px[(kTH - 6) * (kTW - 4 - 6)];
  __shared__ uint32_t s_list[kCap];
  __shared__ int s_n, s_base, s_npx;

  const int4 ce = __ldg(cells + blockIdx.x);
  const int level = ce.x & 0xff, ci = (ce.x >> 8) & 0xfff, cj = (ce.x >> 20) & 0xfff;
  const int x0 = ce.y & 0xffff, y0 = ce.y >> 16, cw = ce.z & 0xffff, ch = ce.z >> 16;
  const LevelGeom& L = g.lv[level];
  const int pitch = L.pitch;
  const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const uint8_t* plane = pyr + (long long)f * g.frame_bytes + L.plane_off;
  const long long org = (long long)(y0 + kBorder) * pitch + kXOff + x0;
  const int shift = (int)(org & 3);
  const int words = (shift + cw + 3) >> 2;
  const uint8_t* src = plane + org - shift;
  if (kTW <= 64) {
    // rows of at most 16 words: a half warp per row, two rows per warp and step — no index division at all
    const int w = lane & 15;
    for (int r = (tid >> 4); r < ch; r += kThreads / 16)
      if (w < words) *(uint32_t*)(tile + r * kTW + 4 * w) = __ldg((const uint32_t*)(src + (long long)r * pitch) + w);
  } else {
    const unsigned wmagic = kMagic20[words];      // i / words == (i * wmagic) >> 20 for i < 2^13
    for (int i = tid; i < ch * words; i += kThreads) {
      int r = (int)(((unsigned)i * wmagic) >> 20), w = i - r * words;
      uint32_t v = __ldg((const uint32_t*)(src + (long long)r * pitch) + w);
      *(uint32_t*)(tile + r * kTW + 4 * w) = v;
    }
  }
  for (int i = tid; i < ch * (kTW / 16); i += kThreads) ((uint4*)score)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { s_n = 0; s_npx = 0; }
  __syncthreads();

  const int dw = cw - 6, dh = ch - 6;   // detection area
  const int npx = dw > 0 && dh > 0 ? dw * dh : 0;
  int th = g.ini_th;
  for (int pass = 0; pass < 2; pass++) {
    // A: compass filter on four pixels at a time (one aligned tile word), compaction.  Only a NECESSARY condition is
    // needed here (B is exact): |ring - c| > t on (0 or 8) and on (4 or 12), polarity ignored, thresholded on
    // |d| >> 1 so that the test is an add into bit 7 of every byte.
    {
      const int c0 = shift + 3, w_first = c0 >> 2, nw = ((c0 + dw - 1) >> 2) - w_first + 1;
      const int items = npx > 0 ? dh * nw : 0;
      const unsigned nmagic = kMagic20[nw];
      const unsigned full_words_end = dw > 3 ? (unsigned)(dw - 3) : 0u;     // words with 0 <= xb < this have four detection pixels
      const unsigned addc = (0x80u - (unsigned)((th + 1) >> 1)) * 0x01010101u;
      const uint32_t* t32 = (const uint32_t*)tile;
      for (int base = 0; base < items; base += kThreads) {
        const int i = base + tid;
        unsigned hit = 0;
        int y = 0, xb = 0;
        if (i < items) {
          y = (int)(((unsigned)i * nmagic) >> 20);
          const int wc = w_first + (i - y * nw);
          const uint32_t* row = t32 + (y + 3) * (kTW / 4) + wc;
          const uint32_t C = row[0], Lw = row[-1], Rw = row[1];
          const uint32_t U = row[-3 * (kTW / 4)], D = row[3 * (kTW / 4)];
          const unsigned a0 = __vabsdiffu4(D, C), a8 = __vabsdiffu4(U, C);
          const unsigned a4 = __vabsdiffu4(__funnelshift_r(C, Rw, 24), C), a12 = __vabsdiffu4(__funnelshift_r(Lw, C, 8), C);
          const unsigned g0 = ((a0 >> 1) & 0x7f7f7f7fu) + addc, g8 = ((a8 >> 1) & 0x7f7f7f7fu) + addc;
          const unsigned g4 = ((a4 >> 1) & 0x7f7f7f7fu) + addc, g12 = ((a12 >> 1) & 0x7f7f7f7fu) + addc;
          hit = (g0 | g8) & (g4 | g12) & 0x80808080u;
          xb = 4 * wc - c0;                       // detection-area x of byte 0 of this word
          // bytes outside [0, dw) are not detection pixels
          if ((unsigned)xb >= full_words_end) {                 // only the first / last word of a row can stick out
#pragma unroll
            for (int j = 0; j < 4; j++)
              if ((unsigned)(xb + j) >= (unsigned)dw) hit &= ~(0x80u << (8 * j));
          }
        }
        if (__any_sync(0xffffffffu, hit != 0)) {
          const unsigned b0 = __ballot_sync(0xffffffffu, hit & 0x80u), b1 = __ballot_sync(0xffffffffu, hit & 0x8000u);
          const unsigned b2 = __ballot_sync(0xffffffffu, hit & 0x800000u), b3 = __ballot_sync(0xffffffffu, hit & 0x80000000u);
          const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
          int off = 0;
          if (lane == 0) off = atomicAdd(&s_npx, n0 + n1 + n2 + n3);
          off = __shfl_sync(0xffffffffu, off, 0);
          const unsigned lt = (1u << lane) - 1;
          const int yx = (y << 8) + xb;           // + j >= y << 8 for every byte that survived the range mask
          if (hit & 0x80u) s_px[off + __popc(b0 & lt)] = (uint16_t)yx;
          if (hit & 0x8000u) s_px[off + n0 + __popc(b1 & lt)] = (uint16_t)(yx + 1);
          if (hit & 0x800000u) s_px[off + n0 + n1 + __popc(b2 & lt)] = (uint16_t)(yx + 2);
          if (hit & 0x80000000u) s_px[off + n0 + n1 + n2 + __popc(b3 & lt)] = (uint16_t)(yx + 3);
        }
      }
    }
    __syncthreads();
    const int nl = s_npx;
    // B: exact scores of the survivors
    for (int j = tid; j < nl; j += kThreads) {
      const int q = s_px[j], y = q >> 8, x = q & 0xff;
      const int m = fast_arc_max(tile + (y + 3) * kTW + shift + x + 3, kTW);
      if (m > th) score[(y + 3) * kTW + x + 3] = (uint8_t)m;
    }
    __syncthreads();
    // C: strict 3x3 non-max suppression (scores <= th are 0 in the map)
    int kept = 0;
    for (int j = tid; j < nl; j += kThreads) {
      const int q = s_px[j], y = q >> 8, x = q & 0xff;
      const uint8_t* s = score + (y + 3) * kTW + x + 3;
      const int m = s[0];
      if (m <= th) continue;
      const int n = max(max(max((int)s[-kTW - 1], (int)s[-kTW]), max((int)s[-kTW + 1], (int)s[-1])),
                        max(max((int)s[1], (int)s[kTW - 1]), max((int)s[kTW], (int)s[kTW + 1])));
      if (n < m) {
        int slot = atomicAdd(&s_n, 1);
        // reference coordinates: cell-local + (j*wCell, i*hCell), relative to minBorder (:820-825)
        if (slot < kCap)
          s_list[slot] = (uint32_t)(x + 3 + cj * L.w_cell) | ((uint32_t)(y + 3 + ci * L.h_cell) << 12) |
                         ((uint32_t)(m - 1) << 24);
        kept++;
      }
    }
    if (__syncthreads_count(kept) > 0 || th == g.min_th) break;
    th = g.min_th;
    if (tid == 0) s_npx = 0;
    __syncthreads();
  }
  if (dbg && (int)blockIdx.x == dbg_cell && f == 0) {   // verification tap
    for (int i = tid; i < kTH * kTW; i += kThreads) { dbg[i] = tile[i]; dbg[kTH * kTW + i] = score[i]; }
    if (tid == 0) { int* q = (int*)(dbg + 2 * kTH * kTW); q[0] = x0; q[1] = y0; q[2] = cw; q[3] = ch; q[4] = shift; q[5] = level; q[6] = g.ini_th; q[7] = g.min_th; }
  }
  const int n = s_n;
  if (n == 0) return;
  if (tid == 0) s_base = atomicAdd(cand_count + f * kMaxLevels + level, n);
  __syncthreads();
  const int base = s_base;
  uint32_t* out = cand + (long long)f * g.cand_frame + L.cand_off;
  if (n > kCap || base + n > L.cand_cap) {
    if (tid == 0) atomicExch(overflow, 1);
    return;
  }
  for (int i = tid; i < n; i += kThreads) out[base + i] = s_list[i];
}

// -------------------------------------------------------------------------------------------------
// Quadtree distribution (DistributeOctTree / DivideNode, ORBextractor.cc:481-763), one CTA per
// (level, frame).  The std::list is emulated by position: every round rebuilds the node array in list
// order (new children reversed at the front, survivors behind), and every candidate carries the list
// position of its node.  Rounds: split all multi-point nodes in list order ("full"), or, once
// size + 3*expandable > N, in (size desc, newest first) order stopping at N ("careful", :673-738).
struct OctNode { short x0, x1, y0, y1; };

__device__ __forceinline__ int oct_quadrant(const OctNode& nd, int x, int y) {
  int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);   // ceil(float(w)/2), :483-484
  int my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
  return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

// exclusive scan of v[0..n) in shared memory, in place; returns the total. All threads must call.
__device__ int block_exscan(int* v, int n, int* warp_tmp) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int per = (n + T - 1) / T;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; i++) sum += v[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) warp_tmp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int nw = T >> 5;
    int w = tid < nw ? warp_tmp[tid] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (tid >= o) wi += t;
    }
    warp_tmp[tid] = wi - w;            // exclusive warp offsets
    if (tid == 31) warp_tmp[32] = wi;  // total (nw <= 32)
  }
  __syncthreads();
  int run = warp_tmp[tid >> 5] + incl - sum;
  const int total = warp_tmp[32];
  for (int i = lo; i < hi; i++) {
    int t = v[i];
    v[i] = run;
    run += t;
  }
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(kOctThreads) k_octree(OrbGeom g, const uint32_t* __restrict__ cand,
                                                       uint16_t* __restrict__ pnode,
                                                       const int* __restrict__ cand_count,
                                                       uint32_t* __restrict__ stage,
                                                       int* __restrict__ level_counts, int maxn) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve
  OctNode* nodeA = (OctNode*)smem_raw;
  OctNode* nodeB = nodeA + maxn;
  int* cntA = (int*)(nodeB + maxn);
  int* cntB = cntA + maxn;
  int* child = cntB + maxn;          // [maxn][4]
  int* rank = child + 4 * maxn;      // processing rank of a todo node, -1 otherwise
  int* proc = rank + maxn;           // node position at processing rank r
  int* scanA = proc + maxn;          // scratch scans
  int* scanB = scanA + maxn;
  unsigned long long* best = (unsigned long long*)(scanB + maxn);   // [maxn]
  __shared__ int warp_tmp[33];
  __shared__ int s_k, s_kp, s_nexp;

  const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, T = kOctThreads;
  const LevelGeom& L = g.lv[level];
  const int N = L.quota;
  const int np = min(cand_count[f * kMaxLevels + level], L.cand_cap);
  const uint32_t* pts = cand + (long long)f * g.cand_frame + L.cand_off;
  uint16_t* pn = pnode + (long long)f * g.cand_frame + L.cand_off;
  int* out_count = level_counts + f * kMaxLevels + level;
  if (np == 0 || L.n_ini < 1) {
    if (tid == 0) *out_count = 0;
    return;
  }

  // ---- roots (:543-570): nIni nodes of width hX; point -> root by float division ----
  const int n_ini = L.n_ini;
  for (int i = tid; i < n_ini; i += T) {
    OctNode nd;
    nd.x0 = (short)(int)(L.hx * (float)i);
    nd.x1 = (short)(int)(L.hx * (float)(i + 1));
    nd.y0 = 0;
    nd.y1 = (short)L.oh;
    nodeA[i] = nd;
    cntA[i] = 0;
  }
  __syncthreads();
  for (int p = tid; p < np; p += T) {
    int x = pts[p] & 0xfff;
    int r = (int)((float)x / L.hx);
    r = min(r, n_ini - 1);
    pn[p] = (uint16_t)r;
    atomicAdd(&cntA[r], 1);
  }
  __syncthreads();
  // drop empty roots (:574-585), keep order
  for (int i = tid; i < n_ini; i += T) scanA[i] = cntA[i] > 0;
  __syncthreads();
  int n = block_exscan(scanA, n_ini, warp_tmp);
  if (n != n_ini) {
    for (int i = tid; i < n_ini; i += T)
      if (cntA[i] > 0) { nodeB[scanA[i]] = nodeA[i]; cntB[scanA[i]] = cntA[i]; }
    for (int p = tid; p < np; p += T) pn[p] = (uint16_t)scanA[pn[p]];
    __syncthreads();
    OctNode* tn = nodeA; nodeA = nodeB; nodeB = tn;
    int* tc = cntA; cntA = cntB; cntB = tc;
  }
  __syncthreads();

  bool careful = false;
  for (;;) {
    // ---- todo set = nodes with more than one point ----
    for (int i = tid; i < n; i += T) { scanA[i] = cntA[i] > 1; rank[i] = -1; }
    __syncthreads();
    if (!careful) {
      // processing order = list order
      for (int i = tid; i < n; i += T) scanB[i] = scanA[i];
      __syncthreads();
      int k = block_exscan(scanB, n, warp_tmp);
      for (int i = tid; i < n; i += T)
        if (scanA[i]) { rank[i] = scanB[i]; proc[scanB[i]] = i; }
      if (tid == 0) s_k = k;
    } else {
      // processing order = (size desc, newest first == position asc): rank by counting
      for (int i = tid; i < n; i += T) scanB[i] = scanA[i];
      __syncthreads();
      int k = block_exscan(scanB, n, warp_tmp);   // compact todo list into proc[] first
      for (int i = tid; i < n; i += T)
        if (scanA[i]) proc[scanB[i]] = i;
      __syncthreads();
      for (int i = tid; i < k; i += T) scanB[i] = proc[i];   // unsorted todo positions
      __syncthreads();
      for (int i = tid; i < k; i += T) {
        int a = scanB[i], ca = cntA[a], r = 0;
        for (int j = 0; j < k; j++) {
          int b = scanB[j], cb = cntA[b];
          r += (cb > ca) || (cb == ca && b < a);
        }
        rank[a] = r;
      }
      __syncthreads();
      for (int i = tid; i < k; i += T) proc[rank[scanB[i]]] = scanB[i];
      if (tid == 0) s_k = k;
    }
    __syncthreads();
    const int k = s_k;
    if (k == 0) break;   // nothing left to split: list size cannot change (:669)

    // ---- child sizes of every todo node ----
    for (int i = tid; i < 4 * n; i += T) child[i] = 0;
    __syncthreads();
    for (int p = tid; p < np; p += T) {
      int a = pn[p];
      if (rank[a] >= 0) {
        uint32_t v = pts[p];
        atomicAdd(&child[4 * a + oct_quadrant(nodeA[a], v & 0xfff, (v >> 12) & 0xfff)], 1);
      }
    }
    __syncthreads();
    // gain per processing rank, inclusive running size
    for (int r = tid; r < k; r += T) {
      int a = proc[r];
      int nch = (child[4 * a] > 0) + (child[4 * a + 1] > 0) + (child[4 * a + 2] > 0) + (child[4 * a + 3] > 0);
      scanA[r] = nch;          // children created by rank r
      scanB[r] = nch - 1;      // list growth
    }
    __syncthreads();
    block_exscan(scanB, k, warp_tmp);   // scanB[r] = growth before rank r
    if (tid == 0) s_kp = k;
    __syncthreads();
    if (careful) {
      // first rank after which size >= N (:731-732): size_after(r) = n + scanB[r] + (nch_r - 1)
      for (int r = tid; r < k; r += T)
        if (n + scanB[r] + scanA[r] - 1 >= N) atomicMin(&s_kp, r + 1);
      __syncthreads();
    }
    const int kp = s_kp;   // ranks [0,kp) are split this round
    for (int r = tid; r < k; r += T)
      if (r >= kp) { rank[proc[r]] = -1; scanA[r] = 0; }
    __syncthreads();
    const int C = block_exscan(scanA, k, warp_tmp);   // scanA[r] = creation index of rank r's first child
    // survivors keep their relative order behind the C new children
    int* keep = scanB;
    for (int i = tid; i < n; i += T) keep[i] = rank[i] < 0;
    __syncthreads();
    block_exscan(keep, n, warp_tmp);
    const int n_new = C + n - kp;
    if (tid == 0) s_nexp = 0;
    __syncthreads();
    // ---- build the new list ----
    for (int i = tid; i < n; i += T) {
      const OctNode nd = nodeA[i];
      if (rank[i] < 0) {
        nodeB[C + keep[i]] = nd;
        cntB[C + keep[i]] = cntA[i];
      } else {
        int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
        int cidx = scanA[rank[i]], nexp = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          int c = child[4 * i + q];
          if (c == 0) continue;
          OctNode ch;
          ch.x0 = (q & 1) ? (short)mx : nd.x0;
          ch.x1 = (q & 1) ? nd.x1 : (short)mx;
          ch.y0 = (q & 2) ? (short)my : nd.y0;
          ch.y1 = (q & 2) ? nd.y1 : (short)my;
          int pos = C - 1 - cidx;   // push_front: later children sit nearer the front
          nodeB[pos] = ch;
          cntB[pos] = c;
          child[4 * i + q] = -(pos + 1);   // remember where the child went (negative marks "position")
          nexp += c > 1;
          cidx++;
        }
        if (nexp) atomicAdd(&s_nexp, nexp);
      }
    }
    __syncthreads();
    for (int p = tid; p < np; p += T) {
      int a = pn[p];
      if (rank[a] < 0) {
        pn[p] = (uint16_t)(C + keep[a]);
      } else {
        uint32_t v = pts[p];
        int q = oct_quadrant(nodeA[a], v & 0xfff, (v >> 12) & 0xfff);
        pn[p] = (uint16_t)(-child[4 * a + q] - 1);
      }
    }
    __syncthreads();
    { OctNode* tn = nodeA; nodeA = nodeB; nodeB = tn; int* tc = cntA; cntA = cntB; cntB = tc; }
    const int n_prev = n;
    n = n_new;
    const int nexp = s_nexp;
    __syncthreads();
    if (n >= N || n == n_prev) break;                       // :669-672, :734-735
    if (!careful && n + 3 * nexp > N) careful = true;       // :673
  }

  // ---- best response per node, first candidate in reference order wins ties (:744-760) ----
  for (int i = tid; i < n; i += T) best[i] = 0ull;
  __syncthreads();
  for (int p = tid; p < np; p += T) {
    uint32_t v = pts[p];
    int x = v & 0xfff, y = (v >> 12) & 0xfff;
    // vToDistributeKeys order: cell row-major, then row-major inside the cell's detection area
    int cj = (x - 3) / L.w_cell, ci = (y - 3) / L.h_cell;
    uint32_t order = (uint32_t)(((ci * L.n_cols + cj) * 128 + (y - ci * L.h_cell)) * 128 + (x - cj * L.w_cell));
    unsigned long long key = ((unsigned long long)(v >> 24) << 32) | (0xffffffffu - order);
    atomicMax(&best[pn[p]], key);
  }
  __syncthreads();
  uint32_t* out = stage + (long long)f * g.kp_cap + L.kp_off;
  for (int i = tid; i < n; i += T) {
    unsigned long long key = best[i];
    uint32_t order = 0xffffffffu - (uint32_t)key;
    int xl = order & 127, yl = (order >> 7) & 127, cell = order >> 14;
    int ci = cell / L.n_cols, cj = cell - ci * L.n_cols;
    uint32_t x = xl + cj * L.w_cell, y = yl + ci * L.h_cell;
    out[i] = x | (y << 12) | ((uint32_t)(key >> 32) << 24);
  }
  if (tid == 0) *out_count = n;
}

// -------------------------------------------------------------------------------------------------
// cv::GaussianBlur 7x7 sigma 2 (Appendix A.2) of every level's interior; reads the reflect-101 border
// that the pyramid planes already carry (equivalent to blurring the un-bordered clone, :1085-1086).
// Q8 kernel [18,34,48,56,48,34,18]: the row pass is two DP4A per pixel (byte windows cut out of three words with
// funnel shifts); its 16-bit sums are stored as (row 2p, row 2p+1) pairs so the column pass is four DP2A per
// pixel, and one thread produces a 4-column x 2-row block from the same four pair loads.
// ---- TMA helpers: threads arm an mbarrier with the byte count and issue bulk copies, the CTA waits on the barrier -------
// Only the descriptor-less bulk copy (cp.async.bulk, SASS UBLKCP) is used: on this pool's B200s every tensor-map copy
// (cp.async.bulk.tensor / UTMALDG, also the CUDA programming guide's own cde:: example) raises "illegal instruction"
// although cuTensorMapEncodeTiled succeeds — tools/micro/tma.cu, tma2.cu.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {   // 16-byte aligned
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// kTma: every row of the (TH+6)-row input tile arrives by one 160-byte bulk copy (TMA engine, completion on an mbarrier)
// from the 16-byte aligned column x0 - 16 instead of 5-6 bounds-checked word loads per thread; the arithmetic is the same.
template <bool kTma>
__global__ void __launch_bounds__(256) k_blur(OrbGeom g, const int2* __restrict__ tiles,
                                              const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
  constexpr int kPairs = (kBlurTH + 6) / 2;                          // 19 row pairs = tile rows y0-3 .. y0+TH+2
  constexpr int kPitch = kTma ? kBlurTW + 32 : kBlurInPitch;         // bytes per staged row
  constexpr int kLead = kTma ? 12 : 0;                               // staged byte of column x0 - 4
  static_assert(kBlurTH % 2 == 0 && kBlurTW % 4 == 0 && kPitch % 16 == 0, "tile shape");
  __shared__ __align__(128) uint8_t in[(kBlurTH + 6) * kPitch];
  __shared__ __align__(16) uint32_t hp[kPairs * kBlurTW];            // (h[2p][x] | h[2p+1][x] << 16)
  __shared__ __align__(8) uint64_t s_bar;
  const int2 te = __ldg(tiles + blockIdx.x);
  const int level = te.x & 0xff, x0 = (te.x >> 8) * kBlurTW, y0 = te.y * kBlurTH;
  const LevelGeom& L = g.lv[level];
  const int f = blockIdx.y, tid = threadIdx.x;
  const uint8_t* plane = pyr + (long long)f * g.frame_bytes + L.plane_off;
  // input rows y0-3 .. y0+TH+2, columns x0-4 .. x0+TW+3 (word aligned: kXOff + x0 - 4 is a multiple of 4)
  if (kTma) {
    const int row0 = y0 - 3 + kBorder;
    const int n_valid = min(max(L.rows - row0, 0), kBlurTH + 6);    // rows below the plane read as zero
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&s_bar, (unsigned)n_valid * kPitch);
    if (tid < kBlurTH + 6) {
      if (tid < n_valid) bulk_load(in + tid * kPitch, plane + (long long)(row0 + tid) * L.pitch + (kXOff + x0 - 16), kPitch, &s_bar);
      else
        for (int w = 0; w < kPitch / 4; w++) ((uint32_t*)(in + tid * kPitch))[w] = 0;
    }
    __syncthreads();                 // zero rows written
    mbar_wait(&s_bar, 0);
  } else {
    constexpr int in_words = kBlurInPitch / 4;
    for (int i = tid; i < (kBlurTH + 6) * in_words; i += 256) {
      int r = i / in_words, w = i - r * in_words;
      int row = y0 - 3 + r + kBorder, col = kXOff + x0 - 4 + 4 * w;
      uint32_t v = 0;
      if (row < L.rows && col + 3 < L.pitch) v = __ldg((const uint32_t*)(plane + (long long)row * L.pitch + col));
      *(uint32_t*)(in + r * kBlurInPitch + 4 * w) = v;
    }
    __syncthreads();
  }
  constexpr unsigned kWA = 0x38302212u, kWB = 0x00122230u;           // taps 0..3 and 4..6 (+ a zero)
  for (int i = tid; i < kPairs * (kBlurTW / 4); i += 256) {
    const int p = i / (kBlurTW / 4), x4 = (i - p * (kBlurTW / 4)) * 4;
    uint32_t h[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const uint32_t* s = (const uint32_t*)(in + (2 * p + rr) * kPitch + kLead + x4);   // s[0] byte 1 is pixel x4-3
      const uint32_t w0 = s[0], w1 = s[1], w2 = s[2];
      h[rr][0] = __dp4a(__funnelshift_r(w0, w1, 8), kWA, __dp4a(__funnelshift_r(w1, w2, 8), kWB, 0u));
      h[rr][1] = __dp4a(__funnelshift_r(w0, w1, 16), kWA, __dp4a(__funnelshift_r(w1, w2, 16), kWB, 0u));
      h[rr][2] = __dp4a(__funnelshift_r(w0, w1, 24), kWA, __dp4a(__funnelshift_r(w1, w2, 24), kWB, 0u));
      h[rr][3] = __dp4a(w1, kWA, __dp4a(w2, kWB, 0u));
    }
    *(uint4*)(hp + p * kBlurTW + x4) = make_uint4(h[0][0] | (h[1][0] << 16), h[0][1] | (h[1][1] << 16),
                                                  h[0][2] | (h[1][2] << 16), h[0][3] | (h[1][3] << 16));
  }
  __syncthreads();
  uint8_t* dst = blur + (long long)f * g.frame_bytes + L.plane_off;
  // output rows r = 2j (taps = tile rows r..r+6) and r+1 (taps r+1..r+7) read the same pairs j..j+3
  constexpr unsigned kE01 = 0x2212u, kE23 = 0x3830u, kE45 = 0x2230u, kE67 = 0x0012u;    // even row: 18,34 | 48,56 | 48,34 | 18,0
  constexpr unsigned kO01 = 0x1200u, kO23 = 0x3022u, kO45 = 0x3038u, kO67 = 0x1222u;    // odd row:  0,18 | 34,48 | 56,48 | 34,18
  for (int i = tid; i < (kBlurTH / 2) * (kBlurTW / 4); i += 256) {
    const int j = i / (kBlurTW / 4), x4 = (i - j * (kBlurTW / 4)) * 4;
    const int y = y0 + 2 * j, x = x0 + x4;
    if (y >= L.h || x >= L.w) continue;
    const uint4 a = *(const uint4*)(hp + j * kBlurTW + x4), b = *(const uint4*)(hp + (j + 1) * kBlurTW + x4);
    const uint4 c = *(const uint4*)(hp + (j + 2) * kBlurTW + x4), d = *(const uint4*)(hp + (j + 3) * kBlurTW + x4);
    const uint32_t pa[4] = {a.x, a.y, a.z, a.w}, pb[4] = {b.x, b.y, b.z, b.w}, pc[4] = {c.x, c.y, c.z, c.w},
                   pd[4] = {d.x, d.y, d.z, d.w};
    uint32_t even = 0, odd = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t e = __dp2a_lo(pa[k], kE01, __dp2a_lo(pb[k], kE23, __dp2a_lo(pc[k], kE45, __dp2a_lo(pd[k], kE67, 32768u))));
      const uint32_t o = __dp2a_lo(pa[k], kO01, __dp2a_lo(pb[k], kO23, __dp2a_lo(pc[k], kO45, __dp2a_lo(pd[k], kO67, 32768u))));
      even |= (e >> 16) << (8 * k);
      odd |= (o >> 16) << (8 * k);
    }
    // bytes past the interior width land in the border columns of the blurred plane, which nothing reads
    uint8_t* o0 = dst + (long long)(y + kBorder) * L.pitch + kXOff + x;
    *(uint32_t*)o0 = even;
    if (y + 1 < L.h) *(uint32_t*)(o0 + L.pitch) = odd;
  }
}

// -------------------------------------------------------------------------------------------------
// IC_Angle (:77-104) + computeOrbDescriptor (:108-147), one warp per output keypoint; also writes the
// cv::KeyPoint record with pt scaled to level-0 coordinates (:1095-1101).
#ifndef CMOS_DESC_STAGE
#define CMOS_DESC_STAGE 0      // measured on B200: staging 0.260 ms vs direct gathers 0.236 ms per 64 frames — the gathers are not the limit
#endif
constexpr int kWinWords = 10;      // 37 bytes + up to 3 of misalignment
#ifndef CMOS_DESC_MINBLOCKS
#define CMOS_DESC_MINBLOCKS 6      // 40 registers, 48 warps per SM: 0.248 -> 0.236 ms per 64 frames (4: 0.258, 8: 0.248)
#endif
#if CMOS_DESC_MINBLOCKS > 0
__global__ void __launch_bounds__(kDescThreads, CMOS_DESC_MINBLOCKS) k_describe(
#else
__global__ void __launch_bounds__(kDescThreads) k_describe(
#endif
    OrbGeom g, const uint32_t* __restrict__ stage,
                                                          const int* __restrict__ level_counts,
                                                          const uint8_t* __restrict__ pyr,
                                                          const uint8_t* __restrict__ blur,
                                                          const int8_t* __restrict__ pattern,
                                                          cmos_keypoint* __restrict__ kps,
                                                          uint8_t* __restrict__ desc, int* __restrict__ counts) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (kDescThreads / 32) + (threadIdx.x >> 5);
  int level = -1, idx = 0, total = 0;
  for (int l = 0; l < g.nlevels; l++) {
    int c = level_counts[f * kMaxLevels + l];
    if (level < 0 && slot < total + c) { level = l; idx = slot - total; }
    total += c;
  }
  if (slot == 0 && lane == 0) counts[f] = total;
  if (level < 0) return;
  // (the level's geometry is read into registers ONCE: `g.lv[level].pitch` is a dynamically indexed constant-bank load, and
  // the compiler re-issued it — plus the 64-bit address arithmetic behind it — under every predicate of the unrolled loops
  // below: the source-level counters of round 2 showed ~45 % of this kernel's instructions in the 31-row centroid loop)
  const LevelGeom& L = g.lv[level];
  const int pitch = L.pitch;
  const uint32_t v = stage[(long long)f * g.kp_cap + L.kp_off + idx];
  const int x = (int)(v & 0xfff) + kMinBorder, y = (int)((v >> 12) & 0xfff) + kMinBorder;   // level coords
  const long long plane = (long long)f * g.frame_bytes + L.plane_off;
  const uint8_t* c = pyr + plane + (long long)(y + kBorder) * pitch + kXOff + x;

  // intensity centroid over the 31-px disc: lane = column u+15.  A pointer walks down the column; the column sum is
  // multiplied by u once (m10 = u * sum_v I), m01 takes the row number as an immediate.
  // Column u of the disc spans rows -vlim..vlim (umax is non-increasing): rows +r and -r share ONE comparison, the loads
  // are unconditional (the 19-pixel border keeps all 31 rows inside the plane) and masked by a select — 31 predicated
  // loads made the compiler pack 31 long-lived predicates into a bit register and spill.
  int m10 = 0, m01 = 0;
  const int u = lane - 15;
  if (lane < 31) {
    const int vlim = g.vlim[u < 0 ? -u : u];
    const uint8_t* p0 = c + u;
    int colsum = *p0;
#pragma unroll
    for (int r = 1; r <= 15; r++) {
      int ip = p0[(long long)r * pitch], im = p0[-(long long)r * pitch];
      const bool in = r <= vlim;
      ip = in ? ip : 0; im = in ? im : 0;
      colsum += ip + im;
      m01 += r * (ip - im);
    }
    m10 = u * colsum;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  // (measured and dropped: handing the moments over in shared memory so that ONE thread per keypoint computes angle, cosine and
  // sine for the CTA's eight keypoints — the scheme that took a quarter off k_sf_lists — costs two CTA barriers here and made
  // this kernel slower, 0.149 -> 0.159 ms: its warps wait on L2 gathers, and the barriers tie eight of them together)
  const float angle = dev_fast_atan2((float)m01, (float)m10);

  // steered BRIEF: lane = descriptor byte
  const float rad = angle * kFactorPi;
  const float a = glibc_cosf(rad), b = glibc_sinf(rad);
  const int8_t* pat = pattern + lane * 32;
  int byte = 0;
#if CMOS_DESC_STAGE
  // The 512 samples of a keypoint lie within 18 px of it: the warp copies that 37 x 37 window of the blurred level into
  // shared memory with aligned word loads (10 words per row, 12 coalesced rounds) and gathers from there — a quarter of
  // the L1 wavefronts of 16 scattered byte loads per lane.
  __shared__ uint32_t s_win[kDescThreads / 32][37 * kWinWords];
  uint32_t* win = s_win[threadIdx.x >> 5];
  {
    const long long row0 = plane + (long long)(y + kBorder - 18) * L.pitch + kXOff + (x - 18);
    const int off = (int)(row0 & 3);
    const uint8_t* base = blur + (row0 - off);
    for (int idx = lane; idx < 37 * kWinWords; idx += 32) {
      const int r = idx / kWinWords, w = idx - r * kWinWords;
      win[idx] = *(const uint32_t*)(base + (long long)r * L.pitch + 4 * w);
    }
    __syncwarp();
    const uint8_t* wb = (const uint8_t*)win + 18 * (4 * kWinWords) + 18 + off;      // the keypoint inside the window
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
      int t0 = wb[__float2int_rn(x0 * b + y0 * a) * (4 * kWinWords) + __float2int_rn(x0 * a - y0 * b)];
      int t1 = wb[__float2int_rn(x1 * b + y1 * a) * (4 * kWinWords) + __float2int_rn(x1 * a - y1 * b)];
      byte |= (t0 < t1) << k;
    }
  }
#else
  const uint8_t* bc = blur + plane + (long long)(y + kBorder) * pitch + kXOff + x;
  // the lane's 16 pattern points arrive as eight aligned words (4 signed bytes each) instead of 32 byte loads
  const uint32_t* pw = (const uint32_t*)pat;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint32_t w = __ldg(pw + k);
    const float x0 = (float)(int8_t)(w & 0xff), y0 = (float)(int8_t)((w >> 8) & 0xff);
    const float x1 = (float)(int8_t)((w >> 16) & 0xff), y1 = (float)(int8_t)(w >> 24);
    const int t0 = bc[__float2int_rn(x0 * b + y0 * a) * pitch + __float2int_rn(x0 * a - y0 * b)];
    const int t1 = bc[__float2int_rn(x1 * b + y1 * a) * pitch + __float2int_rn(x1 * a - y1 * b)];
    byte |= (t0 < t1) << k;
  }
#endif
  desc[((long long)f * g.kp_cap + slot) * 32 + lane] = (uint8_t)byte;
  if (lane == 0) {
    cmos_keypoint kp;
    float fx = (float)x, fy = (float)y;
    if (level != 0) { fx *= L.scale; fy *= L.scale; }
    kp.x = fx; kp.y = fy; kp.size = (float)L.patch; kp.angle = angle; kp.response = (float)(v >> 24);
    kp.octave = level; kp.class_id = -1;
    kps[(long long)f * g.kp_cap + slot] = kp;
  }
}

// Verification taps for the bit-exact float helpers (tests only).
__global__ void k_debug_sincos(const float* __restrict__ deg, float* __restrict__ c, float* __restrict__ s, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float rad = deg[i] * kFactorPi;
  c[i] = glibc_cosf(rad);
  s[i] = glibc_sinf(rad);
}
__global__ void k_debug_atan2(const float* __restrict__ y, const float* __restrict__ x, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dev_fast_atan2(y[i], x[i]);
}

// =================================================================================================
// Host side
// =================================================================================================

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace cmos

using namespace cmos;

struct cmos_orb {
  cmos_orb_params p{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;            // the blur runs here, beside the quadtree (both only need what FAST / the pyramid left)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = true;                    // CMOS_ORB_NO_OVERLAP=1 keeps every kernel on one stream
  bool blur_tma = true;                   // CMOS_BLUR_NO_TMA=1: word loads instead of bulk copies
  std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
  std::vector<int> quota;
  int umax[16];
  int kp_cap = 0;
  // geometry of the current (width, height); cap_* describe the allocation made for (max_w, max_h)
  int cur_w = -1, cur_h = -1;
  OrbGeom geom{};
  int n_cells = 0, n_tiles = 0, oct_maxn = 0;
  size_t cap_frame_bytes = 0, cap_cand_frame = 0, cap_cells = 0, cap_tiles = 0, cap_tab = 0;
  // device buffers
  uint8_t *d_images = nullptr, *d_pyr = nullptr, *d_blur = nullptr, *d_desc = nullptr;
  uint32_t *d_cand = nullptr, *d_stage = nullptr;
  uint16_t* d_pnode = nullptr;
  int *d_cand_count = nullptr, *d_level_counts = nullptr, *d_counts = nullptr, *d_overflow = nullptr;
  int* h_overflow = nullptr;    // pinned landing slot of d_overflow
  cmos_keypoint* d_kps = nullptr;
  int4* d_cells = nullptr;
  int2* d_tiles = nullptr;
  short4* d_tab = nullptr;      // x tables then y tables of every level
  int8_t* d_pattern = nullptr;
  uint8_t* d_dbg = nullptr;
  int dbg_cell = -1;
  std::vector<int> xtab_off, ytab_off;
  size_t images_cap = 0;
  int last_frames = 0, launches = 0;
  int fast_threads = 128;   // CMOS_FAST_THREADS=256 / 64 select other CTA widths (tuning knob).  64 threads (two warps share a cell's
                            // fixed cost instead of four): k_fast alone 0.252 -> 0.236 ms, but the overlapped step gets slower
                            // (155.8 -> 153.8 Mfeat/s, end to end 148.5 -> 141.1): 23 resident CTAs per SM crowd out the other streams' kernels
  bool fast_small_cells = false;   // every cell of the current geometry fits the 48 x 48 tile (CMOS_FAST_LARGE_TILE=1 disables)
  int resize_rows = 2;      // CMOS_RESIZE_ROWS=1|2|4 output rows per thread of k_resize (tuning knob)
  bool resize_v2 = true;    // k_resize2 + k_borders (CMOS_RESIZE_V1=1 selects round 1's k_resize, A/B runs)
  int resize_th = 64;       // output rows per k_resize2 tile (CMOS_RESIZE_TH=16|32|64): the largest one wanted
  int geo_resize_th = 0;    // ... and the largest one whose shared-memory tile holds the source rows of the current image size
                            // (0: none, or a column group leaves the word window: round 1's k_resize with its generic path)
  bool has_result = false;
  StageTimer timer;
};

namespace {

// Geometry for one image size.  Mirrors ComputePyramid (:1107-1132) sizes and the cell grid of
// ComputeKeyPointsOctTree (:765-806).  Also produces the host tables to upload.
struct GeomBuild {
  OrbGeom g{};
  std::vector<int4> cells;
  std::vector<int2> tiles;
  std::vector<short4> tab;
  std::vector<int> xoff, yoff;
  int oct_maxn = 0;
  // what the resize kernels' fast paths assume, checked on the tables of THIS image size:
  bool window_ok = true;              // every aligned group of four output columns reads source bytes 0..8 of its first word
  int max_src_rows[3] = {0, 0, 0};    // source rows a k_resize2 tile of 16 / 32 / 64 output rows reads at most
};

bool build_geometry(const cmos_orb* h, int w, int ht, GeomBuild* out) {
  const int nl = h->p.nlevels;
  OrbGeom& g = out->g;
  std::memset(&g, 0, sizeof(g));
  g.nlevels = nl;
  g.ini_th = h->p.ini_th_fast;
  g.min_th = h->p.min_th_fast;
  for (int i = 0; i < 16; i++) g.umax[i] = h->umax[i];
  for (int au = 0; au < 16; au++) {       // the disc u <= umax[|v|] read by columns: umax is non-increasing in |v|
    int n = 0;                            // (ORBextractor.cc:455-468), so a column is the contiguous range of rows |v| <= vlim
    while (n < 16 && au <= h->umax[n]) n++;
    g.vlim[au] = n - 1;
  }
  long long plane = 0, cand = 0;
  int kp = 0, maxn = 8;
  out->xoff.assign(nl, 0);
  out->yoff.assign(nl, 0);
  for (int l = 0; l < nl; l++) {
    LevelGeom& L = g.lv[l];
    float s = h->inv_sf[l];
    L.w = cv_round_f((float)w * s);
    L.h = cv_round_f((float)ht * s);
    if (L.w <= kBorder || L.h <= kBorder || L.w > 4000 || L.h > 4000) {
      set_error("level %d size %dx%d unsupported (need > %d and <= 4000 per side)", l, L.w, L.h, kBorder);
      return false;
    }
    L.pitch = round_up(kXOff + L.w + kBorder, 128);
    L.rows = L.h + 2 * kBorder;
    L.plane_off = (int)plane;
    plane += (long long)round_up(L.pitch * L.rows, 256);
    L.scale = h->sf[l];
    L.patch = (int)(31 * h->sf[l]);
    L.quota = h->quota[l];
    L.kp_off = kp;
    kp += L.quota + 3;
    const int min_b = kMinBorder, max_bx = L.w - kMinBorder, max_by = L.h - kMinBorder;
    const float width = (float)(max_bx - min_b), height = (float)(max_by - min_b);
    L.n_cols = (int)(width / 30.f);
    L.n_rows = (int)(height / 30.f);
    L.cand_off = (int)cand;
    L.cand_cap = 0;
    L.n_ini = 0;
    if (L.n_cols >= 1 && L.n_rows >= 1) {
      L.w_cell = (int)std::ceil(width / L.n_cols);
      L.h_cell = (int)std::ceil(height / L.n_rows);
      for (int i = 0; i < L.n_rows; i++) {
        const float ini_y = (float)(min_b + i * L.h_cell);
        float max_y = ini_y + L.h_cell + 6;
        if (ini_y >= max_by - 3) continue;
        if (max_y > max_by) max_y = (float)max_by;
        for (int j = 0; j < L.n_cols; j++) {
          const float ini_x = (float)(min_b + j * L.w_cell);
          float max_x = ini_x + L.w_cell + 6;
          if (ini_x >= max_bx - 6) continue;
          if (max_x > max_bx) max_x = (float)max_bx;
          int cw = (int)max_x - (int)ini_x, ch = (int)max_y - (int)ini_y;
          if (cw > kTileW - 4 || ch > kTileH || i > 0xfff || j > 0xfff) {
            set_error("FAST cell %dx%d exceeds the shared-memory tile", cw, ch);
            return false;
          }
          out->cells.push_back(make_int4(l | (i << 8) | (j << 20), (int)ini_x | ((int)ini_y << 16), cw | (ch << 16), 0));
        }
      }
      // strict 3x3 NMS keeps at most one pixel per 2x2 block of the detection area
      L.cand_cap = ((max_bx - min_b) / 2 + 1) * ((max_by - min_b) / 2 + 1);
      L.ow = max_bx - min_b;
      L.oh = max_by - min_b;
      L.n_ini = (int)std::round(static_cast<float>(L.ow) / L.oh);
      if (L.n_ini >= 1) L.hx = static_cast<float>(L.ow) / L.n_ini;
      maxn = std::max(maxn, std::max(L.quota + 3, 4 * std::max(L.n_ini, 1)));
    }
    cand += round_up(L.cand_cap, 64);
    for (int ty = 0; ty * kBlurTH < L.h; ty++)
      for (int tx = 0; tx * kBlurTW < L.w; tx++) out->tiles.push_back(make_int2(l | (tx << 8), ty));
    // resize tables for level l (from l-1): Appendix A.1
    if (l > 0) {
      const LevelGeom& S = g.lv[l - 1];
      if (out->tab.size() & 1) out->tab.push_back(make_short4(0, 0, 0, 0));   // x tables are read as int4 pairs
      out->xoff[l] = (int)out->tab.size();
      const double scale_x = (double)S.w / L.w, scale_y = (double)S.h / L.h;
      for (int dx = 0; dx < L.w; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
        out->tab.push_back(make_short4((short)cv_round_f((1.f - fx) * 2048.f), (short)cv_round_f(fx * 2048.f), (short)sx, 0));
      }
      // the word-window fast path (three aligned words, funnel shifts by 8 * (offset & 3) out of word pair 0 or 1) holds
      // offsets 0..7 of the first tap: a group of four columns whose last tap starts further right needs the generic path
      for (int x0 = 0; x0 < L.w; x0 += 4) {
        const int s0 = out->tab[out->xoff[l] + x0].z, s3 = out->tab[out->xoff[l] + std::min(x0 + 3, L.w - 1)].z;
        if (s3 - 4 * (s0 >> 2) > 7) out->window_ok = false;
      }
      out->yoff[l] = (int)out->tab.size();
      for (int dy = 0; dy < L.h; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        out->tab.push_back(make_short4((short)sy, (short)cv_round_f((1.f - fy) * 2048.f), (short)cv_round_f(fy * 2048.f), 0));
      }
      for (int k = 0; k < 3; k++) {
        const int th = 16 << k;
        for (int r0 = 0; r0 < L.h; r0 += th) {
          const int r1 = std::min(r0 + th, L.h) - 1;
          const int y_lo = std::min(std::max((int)out->tab[out->yoff[l] + r0].x, 0), S.h - 1);
          const int y_hi = std::min(std::max((int)out->tab[out->yoff[l] + r1].x + 1, 0), S.h - 1);
          out->max_src_rows[k] = std::max(out->max_src_rows[k], y_hi - y_lo + 1);
        }
      }
    }
  }
  g.frame_bytes = plane;
  g.cand_frame = cand;
  g.kp_cap = kp;
  out->oct_maxn = round_up(maxn, 32);
  return true;
}

size_t oct_smem_bytes(int maxn) {
  // nodeA,nodeB (8 B) + cntA,cntB + child[4] + rank + proc + scanA + scanB (4 B each) + best (8 B)
  return (size_t)maxn * (2 * 8 + (2 + 4 + 4) * 4 + 8);
}

int ensure_geometry(cmos_orb* h, int w, int ht) {
  if (h->cur_w == w && h->cur_h == ht) return CMOS_OK;
  GeomBuild gb;
  if (!build_geometry(h, w, ht, &gb)) return CMOS_ERR_ARG;
  if ((size_t)gb.g.frame_bytes > h->cap_frame_bytes || (size_t)gb.g.cand_frame > h->cap_cand_frame ||
      gb.cells.size() > h->cap_cells || gb.tiles.size() > h->cap_tiles || gb.tab.size() > h->cap_tab ||
      oct_smem_bytes(gb.oct_maxn) > 200 * 1024) {
    set_error("image %dx%d does not fit the buffers sized for max %dx%d", w, ht, h->p.max_width, h->p.max_height);
    return CMOS_ERR_ARG;
  }
  CMOS_CUDA_OK(cudaMemcpyAsync(h->d_cells, gb.cells.data(), gb.cells.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  CMOS_CUDA_OK(cudaMemcpyAsync(h->d_tiles, gb.tiles.data(), gb.tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
  CMOS_CUDA_OK(cudaMemcpyAsync(h->d_tab, gb.tab.data(), gb.tab.size() * sizeof(short4), cudaMemcpyHostToDevice, h->stream));
  CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));   // host vectors die with gb
  h->geom = gb.g;
  h->n_cells = (int)gb.cells.size();
  {
    bool small = true;
    for (const int4& ce : gb.cells) {
      const int cw = ce.z & 0xffff, ch = ce.z >> 16;
      if (cw > kSmallTileW - 4 || ch > kSmallTileH) { small = false; break; }
    }
    const char* e = std::getenv("CMOS_FAST_LARGE_TILE");
    h->fast_small_cells = small && !(e && e[0] == '1');
  }
  h->n_tiles = (int)gb.tiles.size();
  h->xtab_off = gb.xoff;
  h->ytab_off = gb.yoff;
  h->geo_resize_th = 0;
  if (gb.window_ok)
    for (int k = 2; k >= 0; k--)
      if ((16 << k) <= h->resize_th && gb.max_src_rows[k] <= rz_max_src_rows(16 << k)) { h->geo_resize_th = 16 << k; break; }
  h->oct_maxn = gb.oct_maxn;
  h->cur_w = w;
  h->cur_h = ht;
  return CMOS_OK;
}

int enqueue_extract(cmos_orb* h, const uint8_t* d_images, long long frame_stride, int pitch, int n_frames,
                    cudaStream_t st) {
  const OrbGeom& g = h->geom;
  int launches = 0;
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_cand_count, 0, (size_t)h->p.max_batch * kMaxLevels * sizeof(int), st));
  CMOS_CUDA_OK(cudaMemsetAsync(h->d_overflow, 0, sizeof(int), st));
  NvtxRange nvtx_extract("cmos.orb.extract");
  h->timer.begin(st);
  {
    NvtxRange nvtx_pyr("cmos.orb.pyramid");
    const LevelGeom& L = g.lv[0];
    dim3 blk(32, 8), grid((L.pitch / 16 + 31) / 32, (L.rows + 7) / 8, n_frames);   // pitch is a multiple of 128
    k_level0<<<grid, blk, 0, st>>>(g, d_images, frame_stride, pitch, h->d_pyr);
    launches++;
  }
  if (h->resize_v2 && h->geo_resize_th > 0) {
    for (int l = 1; l < g.nlevels; l++) {
      const LevelGeom& L = g.lv[l];
      const int th = h->geo_resize_th;
      dim3 grid((L.w + kRzTW - 1) / kRzTW, (L.h + th - 1) / th, n_frames);
      if (th == 64) k_resize2<64><<<grid, 256, 0, st>>>(g, l, h->d_tab + h->xtab_off[l], h->d_tab + h->ytab_off[l], h->d_pyr);
      else if (th == 32) k_resize2<32><<<grid, 256, 0, st>>>(g, l, h->d_tab + h->xtab_off[l], h->d_tab + h->ytab_off[l], h->d_pyr);
      else k_resize2<16><<<grid, 256, 0, st>>>(g, l, h->d_tab + h->xtab_off[l], h->d_tab + h->ytab_off[l], h->d_pyr);
      launches++;
    }
    if (g.nlevels > 1) {
      const int n_side = (g.lv[1].h + kSideRows - 1) / kSideRows;            // level 1 is the tallest of levels 1..
      k_borders<<<dim3(n_side + (2 * kBorder + kBorderRowsPerCta - 1) / kBorderRowsPerCta, g.nlevels - 1, n_frames), 256, 0, st>>>(g, 1, n_side, h->d_pyr);
      launches++;
    }
  } else
  for (int l = 1; l < g.nlevels; l++) {
    const LevelGeom& L = g.lv[l];
    const int rr = h->resize_rows;
    dim3 blk(64, 4);
    dim3 grid((L.pitch / 4 + 63) / 64, (L.rows + 4 * rr - 1) / (4 * rr), n_frames);
    if (rr == 1) k_resize<1><<<grid, blk, 0, st>>>(g, l, h->d_tab + h->xtab_off[l], h->d_tab + h->ytab_off[l], h->d_pyr);
    else if (rr == 2) k_resize<2><<<grid, blk, 0, st>>>(g, l, h->d_tab + h->xtab_off[l], h->d_tab + h->ytab_off[l], h->d_pyr);
    else k_resize<4><<<grid, blk, 0, st>>>(g, l, h->d_tab + h->xtab_off[l], h->d_tab + h->ytab_off[l], h->d_pyr);
    launches++;
  }
  h->timer.mark(st);   // stage 0: pyramid
  if (h->n_cells > 0) {
    NvtxRange nvtx_fast("cmos.orb.fast");
    const dim3 fg(h->n_cells, n_frames);
#define CMOS_FAST_LAUNCH(T, TW, TH, CAP) \
  k_fast<T, TW, TH, CAP><<<fg, T, 0, st>>>(g, h->d_cells, h->d_pyr, h->d_cand, h->d_cand_count, h->d_overflow, h->d_dbg, h->dbg_cell)
    const bool small = h->fast_small_cells && h->dbg_cell < 0;      // the debug dump has the large tile's layout
    if (h->fast_threads == 64 && small) CMOS_FAST_LAUNCH(64, kSmallTileW, kSmallTileH, kSmallListCap);
    else if (h->fast_threads <= 128) { if (small) CMOS_FAST_LAUNCH(128, kSmallTileW, kSmallTileH, kSmallListCap); else CMOS_FAST_LAUNCH(128, kTileW, kTileH, kCellListCap); }
    else { if (small) CMOS_FAST_LAUNCH(256, kSmallTileW, kSmallTileH, kSmallListCap); else CMOS_FAST_LAUNCH(256, kTileW, kTileH, kCellListCap); }
#undef CMOS_FAST_LAUNCH
    launches++;
  }
  h->timer.mark(st);   // stage 1: FAST
  // The quadtree (one CTA per frame and level, sequential rounds, 60 us of latency) and the blur (every SM busy, reads only
  // the pyramid) are independent: the blur goes to a side stream and the two overlap.  In profiling mode stage 2 then is the quadtree with the
  // blur running beside it and stage 3 the part of the blur that was not hidden.
  const bool fork = h->overlap && h->side && h->ev_fork && h->ev_join;
  if (fork) {
    CMOS_CUDA_OK(cudaEventRecord(h->ev_fork, st));
    CMOS_CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (h->blur_tma) k_blur<true><<<dim3(h->n_tiles, n_frames), 256, 0, h->side>>>(g, h->d_tiles, h->d_pyr, h->d_blur);
    else k_blur<false><<<dim3(h->n_tiles, n_frames), 256, 0, h->side>>>(g, h->d_tiles, h->d_pyr, h->d_blur);
    CMOS_CUDA_OK(cudaEventRecord(h->ev_join, h->side));
  }
  k_octree<<<dim3(g.nlevels, n_frames), kOctThreads, oct_smem_bytes(h->oct_maxn), st>>>(
      g, h->d_cand, h->d_pnode, h->d_cand_count, h->d_stage, h->d_level_counts, h->oct_maxn);
  launches++;
  h->timer.mark(st);   // stage 2: quadtree
  if (fork) CMOS_CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
  else if (h->blur_tma) k_blur<true><<<dim3(h->n_tiles, n_frames), 256, 0, st>>>(g, h->d_tiles, h->d_pyr, h->d_blur);
  else k_blur<false><<<dim3(h->n_tiles, n_frames), 256, 0, st>>>(g, h->d_tiles, h->d_pyr, h->d_blur);
  launches++;
  h->timer.mark(st);   // stage 3: blur
  k_describe<<<dim3((g.kp_cap + kDescThreads / 32 - 1) / (kDescThreads / 32), n_frames), kDescThreads, 0, st>>>(
      g, h->d_stage, h->d_level_counts, h->d_pyr, h->d_blur, h->d_pattern, h->d_kps, h->d_desc, h->d_counts);
  launches++;
  h->timer.mark(st);   // stage 4: orientation + descriptors
  CMOS_CUDA_OK(cudaGetLastError());
  h->launches = launches;
  h->last_frames = n_frames;
  h->has_result = true;
  return CMOS_OK;
}

}  // namespace

extern "C" {

int cmos_orb_create(const cmos_orb_params* params, cmos_orb_t* out) {
  CMOS_REQUIRE(params && out, "null argument");
  CMOS_REQUIRE(params->nlevels >= 1 && params->nlevels <= kMaxLevels, "nlevels must be in 1..%d", kMaxLevels);
  CMOS_REQUIRE(params->nfeatures > 0 && params->scale_factor > 1.0f, "bad nfeatures/scale_factor");
  CMOS_REQUIRE(params->max_width > 0 && params->max_height > 0 && params->max_batch > 0, "bad max sizes");
  CMOS_REQUIRE(params->min_th_fast >= 1 && params->ini_th_fast >= params->min_th_fast && params->ini_th_fast < 255,
               "FAST thresholds must satisfy 1 <= min <= ini < 255");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return CMOS_ERR_CUDA;
  }
  CMOS_REQUIRE(params->device >= 0 && params->device < ndev, "device %d out of range", params->device);
  CMOS_CUDA_OK(cudaSetDevice(params->device));
  cmos_orb* h = new cmos_orb();
  h->p = *params;
  h->device = params->device;
  if (const char* e = std::getenv("CMOS_FAST_THREADS")) h->fast_threads = std::atoi(e) == 256 ? 256 : std::atoi(e) == 64 ? 64 : 128;
  if (const char* e = std::getenv("CMOS_RESIZE_V1")) h->resize_v2 = !(e[0] == '1');
  if (const char* e = std::getenv("CMOS_RESIZE_TH")) { int t = std::atoi(e); h->resize_th = t == 16 ? 16 : t == 32 ? 32 : 64; }
  if (const char* e = std::getenv("CMOS_RESIZE_ROWS")) { int r = std::atoi(e); h->resize_rows = r == 2 ? 2 : r == 4 ? 4 : 1; }
  // tables of the constructor, ORBextractor.cc:410-470 (scaleFactor member is double, ORBextractor.h:95)
  const int nl = params->nlevels;
  const double scale_d = (double)params->scale_factor;
  h->sf.resize(nl); h->inv_sf.resize(nl); h->sigma2.resize(nl); h->inv_sigma2.resize(nl); h->quota.resize(nl);
  h->sf[0] = 1.f; h->sigma2[0] = 1.f;
  for (int i = 1; i < nl; i++) {
    h->sf[i] = (float)(h->sf[i - 1] * scale_d);
    h->sigma2[i] = h->sf[i] * h->sf[i];
  }
  for (int i = 0; i < nl; i++) { h->inv_sf[i] = 1.0f / h->sf[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
  float factor = (float)(1.0f / scale_d);
  float per = params->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; l++) {
    h->quota[l] = cv_round_f(per);
    sum += h->quota[l];
    per *= factor;
  }
  h->quota[nl - 1] = std::max(params->nfeatures - sum, 0);
  {
    int vmax = (int)std::floor(15 * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(15 * std::sqrt(2.f) / 2);
    for (int v = 0; v < 16; v++) h->umax[v] = 0;
    for (int v = 0; v <= vmax; ++v) h->umax[v] = cv_round_d(std::sqrt(225.0 - v * v));
    for (int v = 15, v0 = 0; v >= vmin; --v) {
      while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
      h->umax[v] = v0;
      ++v0;
    }
  }
  GeomBuild gb;
  if (!build_geometry(h, params->max_width, params->max_height, &gb)) { delete h; return CMOS_ERR_ARG; }
  if (oct_smem_bytes(gb.oct_maxn) > 200 * 1024) {
    set_error("nfeatures too large for the quadtree kernel's shared memory");
    delete h;
    return CMOS_ERR_ARG;
  }
  h->kp_cap = gb.g.kp_cap;
  const size_t B = params->max_batch;
  // a little slack so that smaller images with unluckier rounding still fit
  h->cap_frame_bytes = gb.g.frame_bytes + 64 * 1024;
  h->cap_cand_frame = gb.g.cand_frame + 4096;
  h->cap_cells = gb.cells.size() + 256;
  h->cap_tiles = gb.tiles.size() + 256;
  h->cap_tab = gb.tab.size() + 1024;
  h->images_cap = (size_t)params->max_width * params->max_height * B;
  cudaError_t err = cudaSuccess;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) err = cudaErrorUnknown;
  if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess) err = cudaErrorUnknown;
  if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) err = cudaErrorUnknown;
  if (cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) err = cudaErrorUnknown;
  { const char* e = std::getenv("CMOS_ORB_NO_OVERLAP"); h->overlap = !(e && e[0] == '1'); }
  { const char* e = std::getenv("CMOS_BLUR_NO_TMA"); h->blur_tma = !(e && e[0] == '1'); }
  h->d_images = dev_alloc<uint8_t>(h->images_cap, &err);
  h->d_pyr = dev_alloc<uint8_t>(h->cap_frame_bytes * B, &err);
  h->d_blur = dev_alloc<uint8_t>(h->cap_frame_bytes * B, &err);
  h->d_cand = dev_alloc<uint32_t>(h->cap_cand_frame * B, &err);
  h->d_pnode = dev_alloc<uint16_t>(h->cap_cand_frame * B, &err);
  h->d_stage = dev_alloc<uint32_t>((size_t)h->kp_cap * B, &err);
  h->d_kps = dev_alloc<cmos_keypoint>((size_t)h->kp_cap * B, &err);
  h->d_desc = dev_alloc<uint8_t>((size_t)h->kp_cap * B * 32, &err);
  h->d_cand_count = dev_alloc<int>(B * kMaxLevels, &err);
  h->d_level_counts = dev_alloc<int>(B * kMaxLevels, &err);
  h->d_counts = dev_alloc<int>(B, &err);
  h->d_overflow = dev_alloc<int>(1, &err);
  h->d_cells = dev_alloc<int4>(h->cap_cells, &err);
  h->d_tiles = dev_alloc<int2>(h->cap_tiles, &err);
  h->d_tab = dev_alloc<short4>(h->cap_tab, &err);
  h->d_pattern = dev_alloc<int8_t>(1024, &err);
  if (err != cudaSuccess) {
    set_error("device allocation failed: %s", cudaGetErrorString(err));
    cmos_orb_destroy(h);
    return CMOS_ERR_CUDA;
  }
  cudaMemcpy(h->d_pattern, cmos_orb_pattern_xy, 1024, cudaMemcpyHostToDevice);
  cudaMemset(h->d_blur, 0, h->cap_frame_bytes * B);
  cudaMemset(h->d_level_counts, 0, B * kMaxLevels * sizeof(int));
  cudaMemset(h->d_counts, 0, B * sizeof(int));
  cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  CMOS_CUDA_OK(cudaGetLastError());
  *out = h;
  return CMOS_OK;
}

int cmos_orb_destroy(cmos_orb_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->device);
  void* bufs[] = {h->d_images, h->d_pyr, h->d_blur, h->d_cand, h->d_pnode, h->d_stage, h->d_kps, h->d_desc,
                  h->d_cand_count, h->d_level_counts, h->d_counts, h->d_overflow, h->d_cells, h->d_tiles,
                  h->d_tab, h->d_pattern};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (h->h_overflow) cudaFreeHost(h->h_overflow);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->stream) cudaStreamDestroy(h->stream);
  h->timer.destroy();
  delete h;
  return CMOS_OK;
}

int cmos_orb_set_profiling(cmos_orb_t h, int32_t enable) {
  CMOS_REQUIRE(h, "null handle");
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  h->timer.reset();
  h->timer.enabled = enable != 0;
  return CMOS_OK;
}

int cmos_orb_stage_times(cmos_orb_t h, double* ms, int32_t capacity, int32_t* n_stages, int64_t* calls) {
  CMOS_REQUIRE(h && ms && n_stages && calls, "null argument");
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  h->timer.fold();
  *n_stages = h->timer.n_stages;
  *calls = h->timer.calls;
  for (int i = 0; i < capacity && i < h->timer.n_stages; i++) ms[i] = h->timer.total_ms[i];
  return CMOS_OK;
}

int cmos_orb_get_levels(cmos_orb_t h, int32_t* nlevels) {
  CMOS_REQUIRE(h && nlevels, "null argument");
  *nlevels = h->p.nlevels;
  return CMOS_OK;
}

int cmos_orb_get_scale_factors(cmos_orb_t h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  CMOS_REQUIRE(h, "null handle");
  for (int i = 0; i < h->p.nlevels; i++) {
    if (scale) scale[i] = h->sf[i];
    if (inv_scale) inv_scale[i] = h->inv_sf[i];
    if (sigma2) sigma2[i] = h->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
  }
  return CMOS_OK;
}

int cmos_orb_get_features_per_level(cmos_orb_t h, int32_t* quota) {
  CMOS_REQUIRE(h && quota, "null argument");
  for (int i = 0; i < h->p.nlevels; i++) quota[i] = h->quota[i];
  return CMOS_OK;
}

int cmos_orb_keypoint_capacity(cmos_orb_t h, int32_t* cap) {
  CMOS_REQUIRE(h && cap, "null argument");
  *cap = h->kp_cap;
  return CMOS_OK;
}

int cmos_orb_extract_device(cmos_orb_t h, const uint8_t* d_images, int64_t frame_stride, int32_t pitch,
                            int32_t width, int32_t height, int32_t n_frames, void* stream) {
  CMOS_REQUIRE(h && d_images, "null argument");
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->p.max_batch, "n_frames %d outside 1..%d", n_frames, h->p.max_batch);
  CMOS_REQUIRE(width > 0 && height > 0 && width <= h->p.max_width && height <= h->p.max_height && pitch >= width,
               "bad image size %dx%d pitch %d", width, height, pitch);
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  int rc = ensure_geometry(h, width, height);
  if (rc) return rc;
  return enqueue_extract(h, d_images, frame_stride, pitch, n_frames, stream ? (cudaStream_t)stream : h->stream);
}

namespace {
// D2H of the last extraction's results, enqueued on st (no synchronisation).
int enqueue_download(cmos_orb* h, int n_frames, cmos_keypoint* keypoints, uint8_t* descriptors, int* counts,
                     int capacity, cudaStream_t st) {
  if (!h->h_overflow) CMOS_CUDA_OK(cudaHostAlloc((void**)&h->h_overflow, sizeof(int), cudaHostAllocDefault));
  CMOS_CUDA_OK(cudaMemcpyAsync(counts, h->d_counts, n_frames * sizeof(int), cudaMemcpyDeviceToHost, st));
  CMOS_CUDA_OK(cudaMemcpyAsync(h->h_overflow, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  const size_t krow = (size_t)h->kp_cap * sizeof(cmos_keypoint), drow = (size_t)h->kp_cap * 32;
  if (keypoints) {
    if (capacity == h->kp_cap)
      CMOS_CUDA_OK(cudaMemcpyAsync(keypoints, h->d_kps, krow * n_frames, cudaMemcpyDeviceToHost, st));
    else
      CMOS_CUDA_OK(cudaMemcpy2DAsync(keypoints, (size_t)capacity * sizeof(cmos_keypoint), h->d_kps, krow, krow, n_frames,
                                     cudaMemcpyDeviceToHost, st));
  }
  if (descriptors) {
    if (capacity == h->kp_cap)
      CMOS_CUDA_OK(cudaMemcpyAsync(descriptors, h->d_desc, drow * n_frames, cudaMemcpyDeviceToHost, st));
    else
      CMOS_CUDA_OK(cudaMemcpy2DAsync(descriptors, (size_t)capacity * 32, h->d_desc, drow, drow, n_frames,
                                     cudaMemcpyDeviceToHost, st));
  }
  return CMOS_OK;
}
}  // namespace

int cmos_orb_finish(cmos_orb_t h, void* stream) {
  CMOS_REQUIRE(h, "null handle");
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  CMOS_CUDA_OK(cudaStreamSynchronize(stream ? (cudaStream_t)stream : h->stream));
  if (h->h_overflow && *h->h_overflow) {
    *h->h_overflow = 0;
    set_error("FAST candidate buffer overflow");
    return CMOS_ERR_CAPACITY;
  }
  return CMOS_OK;
}

// The overflow word alone (pinned host memory, written behind the kernels): for callers that waited on an event of their own.
int cmos_orb_check_overflow(cmos_orb_t h) {
  CMOS_REQUIRE(h, "null handle");
  if (h->h_overflow && *h->h_overflow) {
    *h->h_overflow = 0;
    set_error("FAST candidate buffer overflow");
    return CMOS_ERR_CAPACITY;
  }
  return CMOS_OK;
}

int cmos_orb_download(cmos_orb_t h, int32_t n_frames, cmos_keypoint* keypoints, uint8_t* descriptors,
                      int32_t* counts, int32_t capacity, void* stream) {
  CMOS_REQUIRE(h && counts, "null argument");
  if (!h->has_result) { set_error("download before extract"); return CMOS_ERR_STATE; }
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->last_frames, "n_frames %d outside 1..%d", n_frames, h->last_frames);
  CMOS_REQUIRE(capacity >= h->kp_cap, "capacity %d < cmos_orb_keypoint_capacity %d", capacity, h->kp_cap);
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  int rc = enqueue_download(h, n_frames, keypoints, descriptors, counts, capacity, st);
  if (rc) return rc;
  return cmos_orb_finish(h, st);
}

int cmos_orb_extract_async(cmos_orb_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                           int32_t height, int32_t n_frames, cmos_keypoint* keypoints, uint8_t* descriptors,
                           int32_t* counts, int32_t capacity, void* stream) {
  CMOS_REQUIRE(h && counts && images, "null argument");
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->p.max_batch, "n_frames %d outside 1..%d", n_frames, h->p.max_batch);
  CMOS_REQUIRE(width > 0 && height > 0 && width <= h->p.max_width && height <= h->p.max_height && pitch >= width,
               "bad image size %dx%d pitch %d", width, height, pitch);
  CMOS_REQUIRE(capacity >= h->kp_cap, "capacity %d < cmos_orb_keypoint_capacity %d", capacity, h->kp_cap);
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const int64_t frame_bytes = (int64_t)width * height;
  if (pitch == width && (n_frames == 1 || frame_stride == frame_bytes)) {
    CMOS_CUDA_OK(cudaMemcpyAsync(h->d_images, images, (size_t)frame_bytes * n_frames, cudaMemcpyHostToDevice, st));
  } else if (n_frames == 1 || frame_stride == (int64_t)pitch * height) {
    CMOS_CUDA_OK(cudaMemcpy2DAsync(h->d_images, width, images, pitch, width, (size_t)height * n_frames,
                                   cudaMemcpyHostToDevice, st));
  } else {
    for (int f = 0; f < n_frames; f++)
      CMOS_CUDA_OK(cudaMemcpy2DAsync(h->d_images + (size_t)f * frame_bytes, width, images + f * frame_stride, pitch,
                                     width, (size_t)height, cudaMemcpyHostToDevice, st));
  }
  int rc = cmos_orb_extract_device(h, h->d_images, frame_bytes, width, width, height, n_frames, st);
  if (rc) return rc;
  return enqueue_download(h, n_frames, keypoints, descriptors, counts, capacity, st);
}

int cmos_orb_extract(cmos_orb_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                     int32_t height, int32_t n_frames, cmos_keypoint* keypoints, uint8_t* descriptors,
                     int32_t* counts, int32_t capacity) {
  CMOS_REQUIRE(h && counts, "null argument");
  CMOS_REQUIRE(n_frames >= 1 && n_frames <= h->p.max_batch, "n_frames %d outside 1..%d", n_frames, h->p.max_batch);
  if (width == 0 || height == 0 || !images) {   // empty image: silent no-op (ORBextractor.cc:1046-1047)
    for (int f = 0; f < n_frames; f++) counts[f] = 0;
    return CMOS_OK;
  }
  int rc = cmos_orb_extract_async(h, images, frame_stride, pitch, width, height, n_frames, keypoints, descriptors, counts,
                                  capacity, h->stream);
  if (rc) return rc;
  return cmos_orb_finish(h, h->stream);
}

int cmos_orb_device_results(cmos_orb_t h, cmos_keypoint** d_keypoints, uint8_t** d_descriptors, int32_t** d_counts,
                            int32_t** d_level_counts, int32_t* capacity) {
  CMOS_REQUIRE(h, "null handle");
  if (d_keypoints) *d_keypoints = h->d_kps;
  if (d_descriptors) *d_descriptors = h->d_desc;
  if (d_counts) *d_counts = h->d_counts;
  if (d_level_counts) *d_level_counts = h->d_level_counts;
  if (capacity) *capacity = h->kp_cap;
  return CMOS_OK;
}

int cmos_orb_level_size(cmos_orb_t h, int32_t level, int32_t* w, int32_t* ht) {
  CMOS_REQUIRE(h && w && ht && level >= 0 && level < h->p.nlevels, "bad argument");
  if (h->cur_w < 0) { set_error("no extraction yet"); return CMOS_ERR_STATE; }
  *w = h->geom.lv[level].w;
  *ht = h->geom.lv[level].h;
  return CMOS_OK;
}

static int debug_plane(cmos_orb_t h, const uint8_t* base, int frame, int level, uint8_t* out, bool bordered) {
  CMOS_REQUIRE(h && out && level >= 0 && level < h->p.nlevels && frame >= 0 && frame < h->p.max_batch, "bad argument");
  if (!h->has_result) { set_error("no extraction yet"); return CMOS_ERR_STATE; }
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  const LevelGeom& L = h->geom.lv[level];
  CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
  const uint8_t* src = base + (size_t)frame * h->geom.frame_bytes + L.plane_off;
  if (bordered)
    CMOS_CUDA_OK(cudaMemcpy2D(out, L.w + 2 * kBorder, src + kXOff - kBorder, L.pitch, L.w + 2 * kBorder, L.rows,
                              cudaMemcpyDeviceToHost));
  else
    CMOS_CUDA_OK(cudaMemcpy2D(out, L.w, src + (size_t)kBorder * L.pitch + kXOff, L.pitch, L.w, L.h,
                              cudaMemcpyDeviceToHost));
  return CMOS_OK;
}

int cmos_orb_debug_level_image(cmos_orb_t h, int32_t frame, int32_t level, uint8_t* out) {
  return debug_plane(h, h ? h->d_pyr : nullptr, frame, level, out, true);
}
int cmos_orb_debug_level_blurred(cmos_orb_t h, int32_t frame, int32_t level, uint8_t* out) {
  return debug_plane(h, h ? h->d_blur : nullptr, frame, level, out, false);
}

int cmos_orb_debug_level_candidates(cmos_orb_t h, int32_t frame, int32_t level, uint32_t* out, int32_t capacity,
                                    int32_t* n) {
  CMOS_REQUIRE(h && n && level >= 0 && level < h->p.nlevels && frame >= 0 && frame < h->p.max_batch, "bad argument");
  if (!h->has_result) { set_error("no extraction yet"); return CMOS_ERR_STATE; }
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
  int cnt = 0;
  CMOS_CUDA_OK(cudaMemcpy(&cnt, h->d_cand_count + frame * kMaxLevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  const LevelGeom& L = h->geom.lv[level];
  cnt = std::min(cnt, L.cand_cap);
  *n = cnt;
  if (out) {
    CMOS_REQUIRE(capacity >= cnt, "capacity %d < %d candidates", capacity, cnt);
    CMOS_CUDA_OK(cudaMemcpy(out, h->d_cand + (size_t)frame * h->geom.cand_frame + L.cand_off, (size_t)cnt * 4,
                            cudaMemcpyDeviceToHost));
  }
  return CMOS_OK;
}

int cmos_orb_debug_fast_cell(cmos_orb_t h, int32_t cell, uint8_t* out) {
  CMOS_REQUIRE(h, "null handle");
  CMOS_CUDA_OK(cudaSetDevice(h->device));
  const size_t bytes = 2 * kTileH * kTileW + 64;
  if (!h->d_dbg) CMOS_CUDA_OK(cudaMalloc(&h->d_dbg, bytes));
  if (out) {
    CMOS_CUDA_OK(cudaStreamSynchronize(h->stream));
    CMOS_CUDA_OK(cudaMemcpy(out, h->d_dbg, bytes, cudaMemcpyDeviceToHost));
  }
  h->dbg_cell = cell;
  return CMOS_OK;
}

int cmos_orb_last_launch_count(cmos_orb_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

int cmos_debug_sincos_deg(const float* deg, float* cos_out, float* sin_out, int32_t n) {
  CMOS_REQUIRE(deg && cos_out && sin_out && n > 0, "bad argument");
  float *d = nullptr, *c = nullptr, *s = nullptr;
  CMOS_CUDA_OK(cudaMalloc(&d, n * 4)); CMOS_CUDA_OK(cudaMalloc(&c, n * 4)); CMOS_CUDA_OK(cudaMalloc(&s, n * 4));
  CMOS_CUDA_OK(cudaMemcpy(d, deg, n * 4, cudaMemcpyHostToDevice));
  k_debug_sincos<<<(n + 255) / 256, 256>>>(d, c, s, n);
  CMOS_CUDA_OK(cudaMemcpy(cos_out, c, n * 4, cudaMemcpyDeviceToHost));
  CMOS_CUDA_OK(cudaMemcpy(sin_out, s, n * 4, cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(c); cudaFree(s);
  return CMOS_OK;
}

int cmos_debug_fast_atan2(const float* y, const float* x, float* out, int32_t n) {
  CMOS_REQUIRE(y && x && out && n > 0, "bad argument");
  float *dy = nullptr, *dx = nullptr, *d = nullptr;
  CMOS_CUDA_OK(cudaMalloc(&dy, n * 4)); CMOS_CUDA_OK(cudaMalloc(&dx, n * 4)); CMOS_CUDA_OK(cudaMalloc(&d, n * 4));
  CMOS_CUDA_OK(cudaMemcpy(dy, y, n * 4, cudaMemcpyHostToDevice));
  CMOS_CUDA_OK(cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice));
  k_debug_atan2<<<(n + 255) / 256, 256>>>(dy, dx, d, n);
  CMOS_CUDA_OK(cudaMemcpy(out, d, n * 4, cudaMemcpyDeviceToHost));
  cudaFree(dy); cudaFree(dx); cudaFree(d);
  return CMOS_OK;
}

}  // extern "C"
