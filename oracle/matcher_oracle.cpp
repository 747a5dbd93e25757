// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.cpp header for the rules).
//
// CPU restatement of the reference's ORBmatcher hot paths over flattened views:
//   Frame::AssignFeaturesToGrid / PosInGrid      /root/reference/src/Frame.cc:158-173,309-320
//   Frame::GetFeaturesInArea                     /root/reference/src/Frame.cc:243-307
//   ORBmatcher::SearchByProjection(F, points)    /root/reference/src/ORBmatcher.cc:42-119
//   ORBmatcher::RadiusByViewingCos               /root/reference/src/ORBmatcher.cc:121-126
//   ORBmatcher::SearchByProjection(cur, last)    /root/reference/src/ORBmatcher.cc:1161-1271
//   ORBmatcher::ComputeThreeMaxima               /root/reference/src/ORBmatcher.cc:1386-1418
//   ORBmatcher::DescriptorDistance               /root/reference/src/ORBmatcher.cc:1422-1437
//   Frame::isInFrustum, MapPoint::PredictScale   /root/reference/src/Frame.cc:191-241, MapPoint.cc:405-420
// All of this is the reference's own code (no third-party arithmetic), integer Hamming distances plus
// float32/float64 geometry.  The reference has no tests for it.  PINNED TO THE REFERENCE ITSELF:
// tests/test_ref_matcher_parity.py runs the reference's own, unmodified src/ORBmatcher.cc, Frame.cc and MapPoint.cc
// (oracle/_ref/libref.so, recipe oracle/ref_shim/Makefile) over real Frame / MapPoint objects — Frame::Frame end to end,
// the grid, GetFeaturesInArea, isInFrustum / PredictScale, both SearchByProjection variants — and requires identical
// results from this restatement; tests/test_oracle_matcher.py adds brute-force cross-checks.  The one modelling choice is
// the association of Eigen's 3-term product sums (Eigen is unvendored; see oracle/ref_shim/Eigen/Core).
// Float policy as orb_oracle.cpp.
//
// Pointer graphs are flattened: a MapPoint* becomes an index; "F.map_points_[idx] &&
// F.map_points_[idx]->Observations() > 0" becomes a per-keypoint `claimed` byte that the call updates as
// it assigns points, using the per-point `has_obs` byte for points it assigns itself.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int kGridCols = 64, kGridRows = 48;   // Frame.h:45-46
constexpr int kThHigh = 100, kThLow = 50, kHistoLength = 30;   // ORBmatcher.cc:35-37

struct KeyPoint { float x, y, size, angle, response; int32_t octave, class_id; };

inline int hamming256(const uint8_t* a, const uint8_t* b) {
  int32_t pa[8], pb[8];
  std::memcpy(pa, a, 32);
  std::memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    unsigned int v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

struct FrameView {
  int n;
  const KeyPoint* kps;
  const uint8_t* desc;
  const int32_t* grid_start;   // [64*48+1], cell = ix*48+iy
  const int32_t* grid_idx;
  float min_x, max_x, min_y, max_y, inv_w, inv_h;
};

// Frame::GetFeaturesInArea
void features_in_area(const FrameView& F, float x, float y, float r, int min_level, int max_level,
                      std::vector<int>& out) {
  out.clear();
  const int min_cx = std::max(0, (int)std::floor((x - F.min_x - r) * F.inv_w));
  if (min_cx >= kGridCols) return;
  const int max_cx = std::min(kGridCols - 1, (int)std::ceil((x - F.min_x + r) * F.inv_w));
  if (max_cx < 0) return;
  const int min_cy = std::max(0, (int)std::floor((y - F.min_y - r) * F.inv_h));
  if (min_cy >= kGridRows) return;
  const int max_cy = std::min(kGridRows - 1, (int)std::ceil((y - F.min_y + r) * F.inv_h));
  if (max_cy < 0) return;
  const bool check_levels = (min_level > 0) || (max_level >= 0);
  for (int ix = min_cx; ix <= max_cx; ix++)
    for (int iy = min_cy; iy <= max_cy; iy++) {
      const int c = ix * kGridRows + iy;
      for (int k = F.grid_start[c]; k < F.grid_start[c + 1]; k++) {
        const int idx = F.grid_idx[k];
        const KeyPoint& kp = F.kps[idx];
        if (check_levels) {
          if (kp.octave < min_level) continue;
          if (max_level >= 0 && kp.octave > max_level) continue;
        }
        const float dx = kp.x - x, dy = kp.y - y;
        if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(idx);
      }
    }
}

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

}  // namespace

extern "C" {

int match_oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

// Frame::AssignFeaturesToGrid as CSR: cell = ix*48 + iy, indices ascending inside a cell.
// Returns the number of keypoints that fell inside the grid.
int match_oracle_build_grid(const void* kps28, int n, float min_x, float min_y, float inv_w, float inv_h,
                            int32_t* grid_start, int32_t* grid_idx) {
  const KeyPoint* kps = (const KeyPoint*)kps28;
  std::vector<std::vector<int>> cells(kGridCols * kGridRows);
  for (int i = 0; i < n; i++) {
    const int px = (int)std::round((kps[i].x - min_x) * inv_w);
    const int py = (int)std::round((kps[i].y - min_y) * inv_h);
    if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
    cells[px * kGridRows + py].push_back(i);
  }
  int pos = 0;
  for (int c = 0; c < kGridCols * kGridRows; c++) {
    grid_start[c] = pos;
    for (int i : cells[c]) grid_idx[pos++] = i;
  }
  grid_start[kGridCols * kGridRows] = pos;
  return pos;
}

int match_oracle_features_in_area(const void* kps28, int n, const int32_t* grid_start, const int32_t* grid_idx,
                                  const float* bounds6, float x, float y, float r, int min_level, int max_level,
                                  int32_t* out) {
  FrameView F{n, (const KeyPoint*)kps28, nullptr, grid_start, grid_idx,
              bounds6[0], bounds6[1], bounds6[2], bounds6[3], bounds6[4], bounds6[5]};
  std::vector<int> v;
  features_in_area(F, x, y, r, min_level, max_level, v);
  for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
  return (int)v.size();
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th)   ORBmatcher.cc:1161-1271
//   bounds6 = {min_x, max_x, min_y, max_y, grid_element_width_inv, grid_element_height_inv}
//   K4 = {fx, fy, cx, cy} (float32, Frame.cc:141-144);  Tcw = row-major 4x4 double (CurrentFrame.Tcw_)
//   last_flags[i]: bit0 = map point present and not outlier, bit1 = that point has Observations() > 0
//   cur_claimed[i2] (in/out): CurrentFrame.map_points_[i2] && Observations() > 0
//   cur_match[i2] (out): index i of the last-frame keypoint whose map point was assigned, or -1
int match_oracle_search_by_projection_frame(
    const void* cur_kps28, const uint8_t* cur_desc, int cur_n, const int32_t* grid_start, const int32_t* grid_idx,
    const float* bounds6, const float* K4, const float* scale_factors, const double* Tcw,
    const void* last_kps28, int last_n, const uint8_t* last_flags, const double* last_xw, const uint8_t* last_mp_desc,
    float th, int check_orientation, uint8_t* cur_claimed, int32_t* cur_match) {
  FrameView F{cur_n, (const KeyPoint*)cur_kps28, cur_desc, grid_start, grid_idx,
              bounds6[0], bounds6[1], bounds6[2], bounds6[3], bounds6[4], bounds6[5]};
  const KeyPoint* last = (const KeyPoint*)last_kps28;
  const float fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  int nmatches = 0;
  std::vector<int> rot_hist[kHistoLength];
  const float factor = 1.0f / kHistoLength;
  for (int i = 0; i < cur_n; i++) cur_match[i] = -1;
  std::vector<int> cand;
  for (int i = 0; i < last_n; i++) {
    if (!(last_flags[i] & 1)) continue;
    const double* X = last_xw + 3 * i;
    const double pcx = (Tcw[0] * X[0] + Tcw[1] * X[1]) + Tcw[2] * X[2] + Tcw[3];
    const double pcy = (Tcw[4] * X[0] + Tcw[5] * X[1]) + Tcw[6] * X[2] + Tcw[7];
    const double pcz = (Tcw[8] * X[0] + Tcw[9] * X[1]) + Tcw[10] * X[2] + Tcw[11];
    const float xc = (float)pcx, yc = (float)pcy;
    const float invzc = (float)(1.0 / pcz);
    if (invzc < 0) continue;
    const float u = fx * xc * invzc + cx;
    const float v = fy * yc * invzc + cy;
    if (u < F.min_x || u > F.max_x) continue;
    if (v < F.min_y || v > F.max_y) continue;
    const int oct = last[i].octave;
    const float radius = th * scale_factors[oct];
    features_in_area(F, u, v, radius, oct - 1, oct + 1, cand);
    if (cand.empty()) continue;
    const uint8_t* dmp = last_mp_desc + 32 * (size_t)i;
    int best = 256, best_idx = -1;
    for (int i2 : cand) {
      if (cur_claimed[i2]) continue;
      const int d = hamming256(dmp, cur_desc + 32 * (size_t)i2);
      if (d < best) { best = d; best_idx = i2; }
    }
    if (best <= kThHigh) {
      cur_match[best_idx] = i;
      cur_claimed[best_idx] = (last_flags[i] >> 1) & 1;
      nmatches++;
      if (check_orientation) {
        float rot = last[i].angle - F.kps[best_idx].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == kHistoLength) bin = 0;
        rot_hist[bin].push_back(best_idx);
      }
    }
  }
  if (check_orientation) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rot_hist, kHistoLength, i1, i2, i3);
    for (int b = 0; b < kHistoLength; b++)
      if (b != i1 && b != i2 && b != i3)
        for (int idx : rot_hist[b]) { cur_match[idx] = -1; cur_claimed[idx] = 0; nmatches--; }   // :1262 map_points_[idx] = nullptr
  }
  return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th)   ORBmatcher.cc:42-119
//   per point: in_view (is_track_in_view_ && !isBad()), level (track_scale_level_), view_cos, proj (x,y),
//   descriptor, has_obs (Observations() > 0).
//   claimed[idx] (in/out) as above; assign[idx] (out): point index newly written to F.map_points_[idx], else -1
int match_oracle_search_by_projection_points(
    const void* kps28, const uint8_t* desc, int n, const int32_t* grid_start, const int32_t* grid_idx,
    const float* bounds6, const float* scale_factors, int n_points, const uint8_t* in_view,
    const int32_t* level, const float* view_cos, const float* proj_xy, const uint8_t* mp_desc,
    const uint8_t* has_obs, float th, float nn_ratio, uint8_t* claimed, int32_t* assign) {
  FrameView F{n, (const KeyPoint*)kps28, desc, grid_start, grid_idx,
              bounds6[0], bounds6[1], bounds6[2], bounds6[3], bounds6[4], bounds6[5]};
  int nmatches = 0;
  const bool b_factor = th != 1.0;
  for (int i = 0; i < n; i++) assign[i] = -1;
  std::vector<int> cand;
  for (int p = 0; p < n_points; p++) {
    if (!in_view[p]) continue;
    const int lvl = level[p];
    float r = (view_cos[p] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos; float vs double literal
    if (b_factor) r *= th;
    features_in_area(F, proj_xy[2 * p], proj_xy[2 * p + 1], r * scale_factors[lvl], lvl - 1, lvl, cand);
    if (cand.empty()) continue;
    const uint8_t* dmp = mp_desc + 32 * (size_t)p;
    int best = 256, best_lvl = -1, best2 = 256, best_lvl2 = -1, best_idx = -1;
    for (int idx : cand) {
      if (claimed[idx]) continue;
      // mvuRight[idx] > 0 never holds in the monocular system (Frame.cc:127): stereo gate omitted
      const int d = hamming256(dmp, desc + 32 * (size_t)idx);
      if (d < best) {
        best2 = best; best = d; best_lvl2 = best_lvl; best_lvl = F.kps[idx].octave; best_idx = idx;
      } else if (d < best2) {
        best_lvl2 = F.kps[idx].octave; best2 = d;
      }
    }
    if (best <= kThHigh) {
      if (best_lvl == best_lvl2 && best > nn_ratio * best2) continue;
      assign[best_idx] = p;
      claimed[best_idx] = has_obs[p];
      nmatches++;
    }
  }
  return nmatches;
}

// Frame::isInFrustum for a batch of points (Frame.cc:191-241) + MapPoint::PredictScale (MapPoint.cc:405-420).
//   pose: Rcw (row-major 3x3 double), tcw, Ow (double 3 each) -> pose15
//   per point: Xw (double3), normal (double3), min/max distance (float, un-scaled: 0.8/1.2 applied here)
//   out: in_view, proj_xy, level, view_cos
void match_oracle_is_in_frustum(const double* pose15, const float* K4, const float* bounds4, float log_scale_factor,
                                int n_levels, float cos_limit, int n_points, const double* xw, const double* normal,
                                const float* min_dist, const float* max_dist, uint8_t* in_view, float* proj_xy,
                                int32_t* level, float* view_cos) {
  const double* R = pose15; const double* t = pose15 + 9; const double* Ow = pose15 + 12;
  const float fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  for (int p = 0; p < n_points; p++) {
    in_view[p] = 0;
    const double* P = xw + 3 * p;
    const double pc[3] = {(R[0] * P[0] + R[1] * P[1]) + R[2] * P[2] + t[0],
                          (R[3] * P[0] + R[4] * P[1]) + R[5] * P[2] + t[1],
                          (R[6] * P[0] + R[7] * P[1]) + R[8] * P[2] + t[2]};
    const float PcX = (float)pc[0], PcY = (float)pc[1], PcZ = (float)pc[2];
    if (PcZ < 0.0f) continue;
    const float invz = 1.0f / PcZ;
    const float u = fx * PcX * invz + cx, v = fy * PcY * invz + cy;
    if (u < bounds4[0] || u > bounds4[1]) continue;
    if (v < bounds4[2] || v > bounds4[3]) continue;
    const float max_d = 1.2f * max_dist[p], min_d = 0.8f * min_dist[p];
    const double PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    const float dist = (float)std::sqrt((PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2]);
    if (dist < min_d || dist > max_d) continue;
    const double* Pn = normal + 3 * p;
    const float vc = (float)(((PO[0] * Pn[0] + PO[1] * Pn[1]) + PO[2] * Pn[2]) / dist);
    if (vc < cos_limit) continue;
    const float ratio = max_dist[p] / dist;
    int ns = (int)std::ceil(std::log(ratio) / log_scale_factor);
    if (ns < 0) ns = 0;
    else if (ns >= n_levels) ns = n_levels - 1;
    in_view[p] = 1;
    proj_xy[2 * p] = u; proj_xy[2 * p + 1] = v;
    level[p] = ns;
    view_cos[p] = vc;
  }
}

// Frame::UndistortKeyPoints (Frame.cc:329-355) = cv::undistortPoints(mat, mat, K_, dist_coef_, cv::Mat(), K_) on the
// keypoint coordinates.  OpenCV (un-vendored, unpinned in the reference; restated from OpenCV 4.13
// modules/calib3d/src/undistort.dispatch.cpp: cvUndistortPointsInternal with the public wrapper's criteria
// TermCriteria(MAX_ITER, 5, 0.01), i.e. exactly five fixed-point iterations, R = I, P = K): all arithmetic in double,
// inputs and outputs float.  PINNED: tests/test_oracle_matcher.py compares it bit for bit with cv2.undistortPoints.
//   K4 = fx, fy, cx, cy (CV_32F in the reference, widened); dist = k1, k2, p1, p2[, k3] (n_dist = 4 or 5)
void frame_oracle_undistort_points(const float* K4, const float* dist, int n_dist, const float* xy_in, int n,
                                   float* xy_out) {
  const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  double k[14] = {0};
  for (int i = 0; i < n_dist && i < 14; i++) k[i] = dist[i];
  const double ifx = 1. / fx, ify = 1. / fy;
  for (int i = 0; i < n; i++) {
    double x = xy_in[2 * i], y = xy_in[2 * i + 1];
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
      const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
    const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
    xy_out[2 * i] = (float)(xx * ww);
    xy_out[2 * i + 1] = (float)(yy * ww);
  }
}

}  // extern "C"
