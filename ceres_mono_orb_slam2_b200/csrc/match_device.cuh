// Device helpers shared by the matcher translation units (match.cu, match_kf.cu): the Frame / KeyFrame grid walk
// of GetFeaturesInArea (Frame.cc:243-307, KeyFrame.cc:575-620) and the 256-bit Hamming distance.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cmos_common.h"

namespace cmos {

constexpr int kCols = CMOS_GRID_COLS, kRows = CMOS_GRID_ROWS, kCells = kCols * kRows;

struct FrameDev {                 // one current frame on the device
  const cmos_keypoint* kps;
  const uint8_t* desc;
  const int* grid_start;
  const int* grid_idx;
  int n;
  // optional: the keypoints in CSR (cell) order, {x, y, octave << 24 | index}: one coalesced 16-byte load per candidate
  // instead of the index load plus three gathers out of the 28-byte keypoint records (null: walk through grid_idx)
  const float4* cells = nullptr;
};

__device__ __forceinline__ int hamming256(const uint32_t a[8], const uint8_t* __restrict__ b) {
  const uint32_t* w = (const uint32_t*)b;
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a[i] ^ __ldg(w + i));
  return d;
}

struct Window { int min_cx, max_cx, min_cy, max_cy; bool ok; };

// cell range of Frame::GetFeaturesInArea (Frame.cc:250-275)
__device__ __forceinline__ Window make_window(const cmos_camera& cam, float x, float y, float r) {
  Window w;
  w.ok = false;
  w.min_cx = max(0, (int)floorf((x - cam.min_x - r) * cam.grid_element_width_inv));
  if (w.min_cx >= kCols) return w;
  w.max_cx = min(kCols - 1, (int)ceilf((x - cam.min_x + r) * cam.grid_element_width_inv));
  if (w.max_cx < 0) return w;
  w.min_cy = max(0, (int)floorf((y - cam.min_y - r) * cam.grid_element_height_inv));
  if (w.min_cy >= kRows) return w;
  w.max_cy = min(kRows - 1, (int)ceilf((y - cam.min_y + r) * cam.grid_element_height_inv));
  if (w.max_cy < 0) return w;
  w.ok = true;
  return w;
}

// Warp-synchronous walk over the candidates of GetFeaturesInArea in the reference's order (ix outer, iy
// inner, insertion order inside a cell — with cell = ix*48+iy one ix column is one contiguous CSR run).
// Calls visit(idx, pass) with 32 consecutive candidates at a time; pass == false for padding lanes and
// for keypoints rejected by the level / window tests (Frame.cc:283-301).
//
// The runs of the window's columns are FLATTENED into one sequence before they are dealt to the lanes: a window is 3-7 columns
// wide with 3-10 keypoints per column, so one batch per column left 70-90 % of the lanes idle and paid the visitor's Hamming /
// ballot / merge code once per column.  Lane c reads the run of column c, an inclusive scan gives every run its place in the
// sequence, candidate t finds its column by comparing with the scanned ends (ncol - 1 shuffles).  Same candidates in the same
// order, only the batch boundaries move — every visitor ranks by (distance, order) with strict comparisons, which does not
// depend on them.  CMOS_WALK_BY_COLUMN restores the per-column batches (A/B).
template <typename Visit>
__device__ __forceinline__ void walk_window(const FrameDev& F, const Window& w, float x, float y, float r,
                                            int min_level, int max_level, int lane, Visit visit) {
  const bool check_levels = (min_level > 0) || (max_level >= 0);
  auto candidate = [&](int kk, bool& pass, int& idx) {
    int oct;
    float kx, ky;
    if (F.cells) {
      const float4 c = __ldg(F.cells + kk);
      const int packed = __float_as_int(c.z);
      idx = packed & 0xffffff; oct = packed >> 24; kx = c.x; ky = c.y;
    } else {
      idx = F.grid_idx[kk];
      const cmos_keypoint* kp = F.kps + idx;
      oct = kp->octave; kx = kp->x; ky = kp->y;
    }
    if (check_levels) {
      if (oct < min_level) pass = false;
      if (max_level >= 0 && oct > max_level) pass = false;
    }
    const float dx = kx - x, dy = ky - y;
    pass = pass && fabsf(dx) < r && fabsf(dy) < r;
  };
#ifdef CMOS_WALK_BY_COLUMN
  for (int ix = w.min_cx; ix <= w.max_cx; ix++) {
    const int k0 = F.grid_start[ix * kRows + w.min_cy], k1 = F.grid_start[ix * kRows + w.max_cy + 1];
    for (int k = k0; k < k1; k += 32) {
      const int kk = k + lane;
      bool pass = kk < k1;
      int idx = 0;
      if (pass) candidate(kk, pass, idx);
      visit(idx, pass);
    }
  }
#else
  for (int cx0 = w.min_cx; cx0 <= w.max_cx; cx0 += 32) {      // 32 columns at a time (the grid has 64)
    const int ncol = min(32, w.max_cx - cx0 + 1);
    int k0 = 0, cnt = 0;
    if (lane < ncol) {
      const int* gs = F.grid_start + (cx0 + lane) * kRows;
      k0 = gs[w.min_cy];
      cnt = gs[w.max_cy + 1] - k0;
    }
    int incl = cnt;                                            // end of this column's run in the flattened sequence
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int shift = k0 - (incl - cnt);                       // CSR position of candidate t of this column = shift + t
    for (int base = 0; base < total; base += 32) {
      const int t = base + lane;
      int c = 0;
      for (int j = 0; j + 1 < ncol; j++) c += (t >= __shfl_sync(0xffffffffu, incl, j)) ? 1 : 0;
      const int kk = __shfl_sync(0xffffffffu, shift, c) + t;
      bool pass = t < total;
      int idx = 0;
      if (pass) candidate(kk, pass, idx);
      visit(idx, pass);
    }
  }
#endif
}

// Frame::AssignFeaturesToGrid for n_frames frames (defined in match.cu): CSR grid, cell = ix*48+iy.
int launch_build_grid(const cmos_camera& cam, const cmos_keypoint* kps, const int* counts, int stride, int max_kp,
                      int* grid_start, int* grid_idx, int n_frames, cudaStream_t st, float4* cells = nullptr);

}  // namespace cmos
