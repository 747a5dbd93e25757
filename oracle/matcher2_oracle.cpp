// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.cpp header for the rules).
//
// CPU restatement of the remaining ORBmatcher searches over flattened views:
//   SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist)   /root/reference/src/ORBmatcher.cc:1273-1384
//   SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)           /root/reference/src/ORBmatcher.cc:258-361
//   SearchByBoW(KeyFrame*, Frame&, matches)                               /root/reference/src/ORBmatcher.cc:151-256
//   SearchByBoW(KeyFrame*, KeyFrame*, matches12)                          /root/reference/src/ORBmatcher.cc:470-580
//   SearchForInitialization                                               /root/reference/src/ORBmatcher.cc:363-468
//   SearchForTriangulation + CheckDistEpipolarLine                        /root/reference/src/ORBmatcher.cc:582-722,128-149
//   Fuse(KeyFrame*, points, th), Fuse(KeyFrame*, Scw, points, th, repl)   /root/reference/src/ORBmatcher.cc:724-842,844-954
//   SearchBySim3                                                          /root/reference/src/ORBmatcher.cc:956-1159
//   KeyFrame::GetFeaturesInArea / IsInImage                               /root/reference/src/KeyFrame.cc:575-626
//   MapPoint::PredictScale, Get{Min,Max}DistanceInvariance                /root/reference/src/MapPoint.cc:380-420
//   MapPoint::ComputeDistinctiveDescriptors, UpdateNormalAndDepth         /root/reference/src/MapPoint.cc:256-315,335-378
// Own code of the reference, integer Hamming + float32/float64 geometry.  The reference has no tests for any of it.
// PINNED TO THE REFERENCE ITSELF: tests/test_ref_matcher_parity.py runs the reference's own, unmodified src/ORBmatcher.cc,
// KeyFrame.cc and MapPoint.cc (oracle/_ref/libref.so, built by oracle/ref_shim/Makefile against stand-ins for the absent
// OpenCV / Eigen / glog headers) over real KeyFrame / MapPoint objects and requires identical matches from this
// restatement; tests/test_oracle_matcher2.py adds hand-built known answers and brute force.
//
// Flattening.  A MapPoint* is an index into the caller's arrays; predicates on the pointer graph
// that do not change during a call are bytes prepared by the caller ("pMP && !pMP->isBad() && !found.count(pMP)").
// The map-mutating functions (Fuse) return DECISIONS: the search for a point never reads what an earlier point's
// Replace / AddObservation wrote (only the skip predicates do, and the host adapter re-evaluates those in order while
// applying the decisions), so "best keypoint per point" is the whole computation.
// Eigen products are evaluated as ((a0*b0 + a1*b1) + a2*b2) (+ t), like the existing matcher oracle.

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int kGridCols = 64, kGridRows = 48;
constexpr int kThHigh = 100, kThLow = 50, kHistoLength = 30;

struct KeyPoint { float x, y, size, angle, response; int32_t octave, class_id; };

inline int hamming256(const uint8_t* a, const uint8_t* b) {
  uint32_t pa[8], pb[8];
  std::memcpy(pa, a, 32);
  std::memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
  return dist;
}

// A Frame or a KeyFrame as the searches read it.  view17 = {min_x, max_x, min_y, max_y, grid_element_width_inv,
// grid_element_height_inv, fx, fy, cx, cy, log_scale_factor, n_levels, 5 spare}; a KeyFrame's bounds are the Frame's
// truncated to int (KeyFrame.h:179-182) — the caller passes them already truncated.
struct View {
  int n;
  const KeyPoint* kps;
  const uint8_t* desc;
  const int32_t* grid_start;
  const int32_t* grid_idx;
  float min_x, max_x, min_y, max_y, inv_w, inv_h, fx, fy, cx, cy, log_sf;
  int n_levels;
  const float* sf;       // scale_factors_
  const float* sigma2;   // level_sigma2s_
};

View make_view(const void* kps28, const uint8_t* desc, int n, const int32_t* gs, const int32_t* gi, const float* v12,
               const float* sf, const float* sigma2) {
  return View{n, (const KeyPoint*)kps28, desc, gs, gi, v12[0], v12[1], v12[2], v12[3], v12[4], v12[5],
              v12[6], v12[7], v12[8], v12[9], v12[10], (int)v12[11], sf, sigma2};
}

// Frame::GetFeaturesInArea (Frame.cc:243-307); with min_level = max_level = -1 it is KeyFrame::GetFeaturesInArea
// (KeyFrame.cc:575-620: no level filter).
void features_in_area(const View& F, float x, float y, float r, int min_level, int max_level, std::vector<int>& out) {
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

inline bool is_in_image(const View& F, float x, float y) {   // KeyFrame::IsInImage
  return x >= F.min_x && x < F.max_x && y >= F.min_y && y < F.max_y;
}

inline int predict_scale(float max_distance, float current_dist, const View& F) {   // MapPoint.cc:390-420
  const float ratio = max_distance / current_dist;
  int n = (int)std::ceil(std::log(ratio) / F.log_sf);
  if (n < 0) n = 0;
  else if (n >= F.n_levels) n = F.n_levels - 1;
  return n;
}

inline void mat3_vec(const double* R, const double* p, const double* t, double* out) {
  for (int i = 0; i < 3; i++) out[i] = (R[3 * i] * p[0] + R[3 * i + 1] * p[1]) + R[3 * i + 2] * p[2] + (t ? t[i] : 0.0);
}
inline double norm3(const double* v) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

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

inline int rot_bin(float a1, float a2) {
  const float factor = 1.0f / kHistoLength;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == kHistoLength) bin = 0;
  return bin;
}

// Scw -> Rcw, tcw, Ow (ORBmatcher.cc:267-272, 853-858)
void decompose_sim3(const double* S, double* R, double* t, double* Ow) {
  const float scw = (float)std::sqrt((S[0] * S[0] + S[1] * S[1]) + S[2] * S[2]);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R[3 * i + j] = S[4 * i + j] / scw;
    t[i] = S[4 * i + 3] / scw;
  }
  for (int i = 0; i < 3; i++) Ow[i] = -((R[i] * t[0] + R[3 + i] * t[1]) + R[6 + i] * t[2]);
}

// The projection + gates shared by Fuse (both), SearchByProjection(KF, Scw): returns false when the point is skipped.
// :745-785 / :867-907 / :287-325
bool project_into_keyframe(const View& F, const double* R, const double* t, const double* Ow, const double* Xw,
                           const double* normal, float min_dist_raw, float max_dist_raw, float* u_out, float* v_out,
                           int* level_out) {
  double pc[3];
  mat3_vec(R, Xw, t, pc);
  if (pc[2] < 0.0) return false;
  const float invz = (float)(1 / pc[2]);
  const float x = (float)(pc[0] * invz);
  const float y = (float)(pc[1] * invz);
  const float u = F.fx * x + F.cx;
  const float v = F.fy * y + F.cy;
  if (!is_in_image(F, u, v)) return false;
  const float max_d = 1.2f * max_dist_raw, min_d = 0.8f * min_dist_raw;
  const double PO[3] = {Xw[0] - Ow[0], Xw[1] - Ow[1], Xw[2] - Ow[2]};
  const float dist = (float)norm3(PO);
  if (dist < min_d || dist > max_d) return false;
  if (((PO[0] * normal[0] + PO[1] * normal[1]) + PO[2] * normal[2]) < 0.5 * dist) return false;
  *u_out = u; *v_out = v;
  *level_out = predict_scale(max_dist_raw, dist, F);
  return true;
}

// Flattened DBoW2::FeatureVector: node ids ascending, start[nn+1], idx[] (feature indices in insertion order).
struct FeatVec { int nn; const int32_t* node; const int32_t* start; const int32_t* idx; };

}  // namespace

extern "C" {

// ---- SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist)  :1273-1384 -------------
//   Tcw: CurrentFrame.Tcw_ row-major 4x4.  Per keyframe keypoint i: kf_valid (pMP && !isBad && !sAlreadyFound.count),
//   Xw, raw min/max distance, MapPoint descriptor, kf_angle (pKF->undistort_keypoints_[i].angle).
//   cur_has_point[i2] in/out (CurrentFrame.map_points_[i2] != NULL), cur_match[i2] out: i or -1.
int match2_oracle_search_by_projection_reloc(
    const void* kps28, const uint8_t* desc, int n, const int32_t* gs, const int32_t* gi, const float* view12,
    const float* sf, const double* Tcw, int n_kf, const uint8_t* kf_valid, const double* kf_xw, const float* kf_min_d,
    const float* kf_max_d, const uint8_t* kf_desc, const float* kf_angle, float th, int orb_dist, int check_ori,
    uint8_t* cur_has_point, int32_t* cur_match) {
  const View F = make_view(kps28, desc, n, gs, gi, view12, sf, nullptr);
  const double R[9] = {Tcw[0], Tcw[1], Tcw[2], Tcw[4], Tcw[5], Tcw[6], Tcw[8], Tcw[9], Tcw[10]};
  const double t[3] = {Tcw[3], Tcw[7], Tcw[11]};
  double Ow[3];
  for (int i = 0; i < 3; i++) Ow[i] = -((R[i] * t[0] + R[3 + i] * t[1]) + R[6 + i] * t[2]);
  std::vector<int> rot_hist[kHistoLength];
  int nmatches = 0;
  for (int i = 0; i < n; i++) cur_match[i] = -1;
  std::vector<int> cand;
  for (int i = 0; i < n_kf; i++) {
    if (!kf_valid[i]) continue;
    const double* X = kf_xw + 3 * i;
    double xc3[3];
    mat3_vec(R, X, t, xc3);
    const float xc = (float)xc3[0], yc = (float)xc3[1];
    const float invzc = (float)(1.0 / xc3[2]);
    const float u = F.fx * xc * invzc + F.cx;
    const float v = F.fy * yc * invzc + F.cy;
    if (u < F.min_x || u > F.max_x) continue;
    if (v < F.min_y || v > F.max_y) continue;
    const double PO[3] = {X[0] - Ow[0], X[1] - Ow[1], X[2] - Ow[2]};
    const float dist3D = (float)norm3(PO);
    const float max_d = 1.2f * kf_max_d[i], min_d = 0.8f * kf_min_d[i];
    if (dist3D < min_d || dist3D > max_d) continue;
    const int lvl = predict_scale(kf_max_d[i], dist3D, F);
    const float radius = th * sf[lvl];
    features_in_area(F, u, v, radius, lvl - 1, lvl + 1, cand);
    if (cand.empty()) continue;
    int best = 256, best_idx = -1;
    for (int i2 : cand) {
      if (cur_has_point[i2]) continue;
      const int d = hamming256(kf_desc + 32 * (size_t)i, desc + 32 * (size_t)i2);
      if (d < best) { best = d; best_idx = i2; }
    }
    if (best <= orb_dist) {
      cur_has_point[best_idx] = 1;
      cur_match[best_idx] = i;
      nmatches++;
      if (check_ori) rot_hist[rot_bin(kf_angle[i], F.kps[best_idx].angle)].push_back(best_idx);
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rot_hist, kHistoLength, i1, i2, i3);
    for (int b = 0; b < kHistoLength; b++)
      if (b != i1 && b != i2 && b != i3)
        for (int idx : rot_hist[b]) { cur_has_point[idx] = 0; cur_match[idx] = -1; nmatches--; }
  }
  return nmatches;
}

// ---- SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th)  :258-361 ------------------------------------
//   pt_skip[p] = pMP->isBad() || spAlreadyFound.count(pMP);  matched[idx] in/out (vpMatched[idx] != NULL);
//   assign[idx] out: p or -1.
int match2_oracle_search_by_projection_sim3(
    const void* kps28, const uint8_t* desc, int n, const int32_t* gs, const int32_t* gi, const float* view12,
    const float* sf, const double* Scw, int n_points, const uint8_t* pt_skip, const double* xw, const double* normal,
    const float* min_d, const float* max_d, const uint8_t* pt_desc, int th, uint8_t* matched, int32_t* assign) {
  const View F = make_view(kps28, desc, n, gs, gi, view12, sf, nullptr);
  double R[9], t[3], Ow[3];
  decompose_sim3(Scw, R, t, Ow);
  int nmatches = 0;
  for (int i = 0; i < n; i++) assign[i] = -1;
  std::vector<int> cand;
  for (int p = 0; p < n_points; p++) {
    if (pt_skip[p]) continue;
    float u, v; int lvl;
    if (!project_into_keyframe(F, R, t, Ow, xw + 3 * p, normal + 3 * p, min_d[p], max_d[p], &u, &v, &lvl)) continue;
    const float radius = th * sf[lvl];
    features_in_area(F, u, v, radius, -1, -1, cand);
    if (cand.empty()) continue;
    int best = 256, best_idx = -1;
    for (int idx : cand) {
      if (matched[idx]) continue;
      const int kl = F.kps[idx].octave;
      if (kl < lvl - 1 || kl > lvl) continue;
      const int d = hamming256(pt_desc + 32 * (size_t)p, desc + 32 * (size_t)idx);
      if (d < best) { best = d; best_idx = idx; }
    }
    if (best <= kThLow) { matched[best_idx] = 1; assign[best_idx] = p; nmatches++; }
  }
  return nmatches;
}

// ---- Fuse(KeyFrame*, vpMapPoints, th) :724-842 (sim3 == 0, pose15 = Rcw, tcw, Ow) and
//      Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) :844-954 (sim3 == 1, pose = Scw row-major 4x4).
//   Returns per point the keypoint the reference would fuse with (best_idx, -1 if none) and its distance; the
//   number of points with a decision is the return value.  pt_skip: see the file header.
int match2_oracle_fuse(const void* kps28, const uint8_t* desc, int n, const int32_t* gs, const int32_t* gi,
                       const float* view12, const float* sf, const float* inv_sigma2, int sim3, const double* pose,
                       int n_points, const uint8_t* pt_skip, const double* xw, const double* normal, const float* min_d,
                       const float* max_d, const uint8_t* pt_desc, float th, int32_t* best_idx, int32_t* best_dist) {
  const View F = make_view(kps28, desc, n, gs, gi, view12, sf, nullptr);
  double R[9], t[3], Ow[3];
  if (sim3) decompose_sim3(pose, R, t, Ow);
  else { std::memcpy(R, pose, 72); std::memcpy(t, pose + 9, 24); std::memcpy(Ow, pose + 12, 24); }
  int nfused = 0;
  std::vector<int> cand;
  for (int p = 0; p < n_points; p++) {
    best_idx[p] = -1; best_dist[p] = 256;
    if (pt_skip[p]) continue;
    float u, v; int lvl;
    if (!project_into_keyframe(F, R, t, Ow, xw + 3 * p, normal + 3 * p, min_d[p], max_d[p], &u, &v, &lvl)) continue;
    const float radius = th * sf[lvl];
    features_in_area(F, u, v, radius, -1, -1, cand);
    if (cand.empty()) continue;
    int best = sim3 ? INT_MAX : 256, bi = -1;
    for (int idx : cand) {
      const KeyPoint& kp = F.kps[idx];
      const int kl = kp.octave;
      if (kl < lvl - 1 || kl > lvl) continue;
      if (!sim3) {
        const float ex = u - kp.x, ey = v - kp.y;
        const float e2 = ex * ex + ey * ey;
        if (e2 * inv_sigma2[kl] > 5.99) continue;     // float product compared with a double literal
      }
      const int d = hamming256(pt_desc + 32 * (size_t)p, desc + 32 * (size_t)idx);
      if (d < best) { best = d; bi = idx; }
    }
    if (best <= kThLow) { best_idx[p] = bi; best_dist[p] = best; nfused++; }
  }
  return nfused;
}

// ---- SearchBySim3 :956-1159 --------------------------------------------------------------------------------------
//   For each keyframe k in {1,2}: mp_valid (pMP && !isBad), already (vbAlreadyMatchedk), Xw, raw min/max distance,
//   MapPoint descriptor, per keypoint.  pose1 / pose2: R (9) + t (3) of the keyframes.  Intrinsics are pKF1's for both
//   directions (:960-963) — view1's fx..cy are used.  match12[i1] out: idx2 or -1.
int match2_oracle_search_by_sim3(
    const void* kps1, const uint8_t* desc1, int n1, const int32_t* gs1, const int32_t* gi1, const float* view1,
    const void* kps2, const uint8_t* desc2, int n2, const int32_t* gs2, const int32_t* gi2, const float* view2,
    const float* sf, const double* pose1, const double* pose2, float s12, const double* R12, const double* t12,
    const uint8_t* valid1, const uint8_t* already1, const double* xw1, const float* min_d1, const float* max_d1,
    const uint8_t* mpdesc1, const uint8_t* valid2, const uint8_t* already2, const double* xw2, const float* min_d2,
    const float* max_d2, const uint8_t* mpdesc2, float th, int32_t* match12) {
  View F1 = make_view(kps1, desc1, n1, gs1, gi1, view1, sf, nullptr);
  View F2 = make_view(kps2, desc2, n2, gs2, gi2, view2, sf, nullptr);
  F2.fx = F1.fx; F2.fy = F1.fy; F2.cx = F1.cx; F2.cy = F1.cy;
  double sR12[9], sR21[9], t21[3];
  const double inv_s = 1.0 / s12;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { sR12[3 * i + j] = s12 * R12[3 * i + j]; sR21[3 * i + j] = inv_s * R12[3 * j + i]; }
  for (int i = 0; i < 3; i++) t21[i] = -((sR21[3 * i] * t12[0] + sR21[3 * i + 1] * t12[1]) + sR21[3 * i + 2] * t12[2]);
  std::vector<int> m1(n1, -1), m2(n2, -1), cand;
  auto direction = [&](int na, const uint8_t* valid, const uint8_t* already, const double* xw, const float* min_d,
                       const float* max_d, const uint8_t* mpdesc, const double* pose_a, const double* sRba,
                       const double* tba, const View& Fb, std::vector<int>& out) {
    for (int i = 0; i < na; i++) {
      if (!valid[i] || already[i]) continue;
      double pa[3], pb[3];
      mat3_vec(pose_a, xw + 3 * i, pose_a + 9, pa);
      mat3_vec(sRba, pa, tba, pb);
      if (pb[2] < 0.0) continue;
      const float invz = (float)(1.0 / pb[2]);
      const float x = (float)(pb[0] * invz), y = (float)(pb[1] * invz);
      const float u = Fb.fx * x + Fb.cx, v = Fb.fy * y + Fb.cy;
      if (!is_in_image(Fb, u, v)) continue;
      const float maxd = 1.2f * max_d[i], mind = 0.8f * min_d[i];
      const float dist3D = (float)norm3(pb);
      if (dist3D < mind || dist3D > maxd) continue;
      const int lvl = predict_scale(max_d[i], dist3D, Fb);
      const float radius = th * sf[lvl];
      features_in_area(Fb, u, v, radius, -1, -1, cand);
      if (cand.empty()) continue;
      int best = INT_MAX, bi = -1;
      for (int idx : cand) {
        const int kl = Fb.kps[idx].octave;
        if (kl < lvl - 1 || kl > lvl) continue;
        const int d = hamming256(mpdesc + 32 * (size_t)i, Fb.desc + 32 * (size_t)idx);
        if (d < best) { best = d; bi = idx; }
      }
      if (best <= kThHigh) out[i] = bi;
    }
  };
  direction(n1, valid1, already1, xw1, min_d1, max_d1, mpdesc1, pose1, sR21, t21, F2, m1);
  direction(n2, valid2, already2, xw2, min_d2, max_d2, mpdesc2, pose2, sR12, t12, F1, m2);
  int found = 0;
  for (int i1 = 0; i1 < n1; i1++) {
    match12[i1] = -1;
    const int idx2 = m1[i1];
    if (idx2 >= 0 && m2[idx2] == i1) { match12[i1] = idx2; found++; }
  }
  return found;
}

// ---- SearchByBoW(KeyFrame*, Frame&) :151-256 (mode 0) and SearchByBoW(KeyFrame*, KeyFrame*) :470-580 (mode 1) -----
//   side 1 = the keyframe whose map points are matched; valid1[i] = pMP && !isBad.  mode 1: valid2 likewise.
//   mode 0: match[idx2] out = idx1 (vpMapPointMatches[realIdxF] = pMP of realIdxKF), accepted when best <= TH_LOW;
//   mode 1: match[idx1] out = idx2, accepted when best < TH_LOW.
int match2_oracle_search_by_bow(int mode, const uint8_t* desc1, const float* angle1, int n1, const uint8_t* valid1,
                                int nn1, const int32_t* node1, const int32_t* start1, const int32_t* idx1,
                                const uint8_t* desc2, const float* angle2, int n2, const uint8_t* valid2, int nn2,
                                const int32_t* node2, const int32_t* start2, const int32_t* idx2, float nn_ratio,
                                int check_ori, int32_t* match) {
  const FeatVec A{nn1, node1, start1, idx1}, B{nn2, node2, start2, idx2};
  std::vector<int> rot_hist[kHistoLength];
  std::vector<uint8_t> taken2(n2, 0);
  const int n_out = mode == 0 ? n2 : n1;
  for (int i = 0; i < n_out; i++) match[i] = -1;
  int nmatches = 0, a = 0, b = 0;
  while (a < A.nn && b < B.nn) {
    if (A.node[a] == B.node[b]) {
      for (int k1 = A.start[a]; k1 < A.start[a + 1]; k1++) {
        const int i1 = A.idx[k1];
        if (!valid1[i1]) continue;
        int best1 = 256, best_i2 = -1, best2 = 256;
        for (int k2 = B.start[b]; k2 < B.start[b + 1]; k2++) {
          const int i2 = B.idx[k2];
          if (taken2[i2]) continue;
          if (mode == 1 && !valid2[i2]) continue;
          const int d = hamming256(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
          if (d < best1) { best2 = best1; best1 = d; best_i2 = i2; }
          else if (d < best2) best2 = d;
        }
        const bool pass = mode == 0 ? best1 <= kThLow : best1 < kThLow;
        if (pass && (float)best1 < nn_ratio * (float)best2) {
          taken2[best_i2] = 1;
          if (mode == 0) match[best_i2] = i1; else match[i1] = best_i2;
          if (check_ori) rot_hist[rot_bin(angle1[i1], angle2[best_i2])].push_back(mode == 0 ? best_i2 : i1);
          nmatches++;
        }
      }
      a++; b++;
    } else if (A.node[a] < B.node[b]) {
      a = (int)(std::lower_bound(A.node, A.node + A.nn, B.node[b]) - A.node);
    } else {
      b = (int)(std::lower_bound(B.node, B.node + B.nn, A.node[a]) - B.node);
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rot_hist, kHistoLength, i1, i2, i3);
    for (int h = 0; h < kHistoLength; h++)
      if (h != i1 && h != i2 && h != i3)
        for (int idx : rot_hist[h]) { match[idx] = -1; nmatches--; }
  }
  return nmatches;
}

// ---- SearchForTriangulation :582-722 (monocular: mvuRight < 0 everywhere, bOnlyStereo false) ----------------------
//   has_point{1,2}[i] = pKF->GetMapPoint(i) != NULL;  epipole (ex, ey) of camera 1 in image 2 (:588-594) is computed
//   here from Cw (camera centre of KF1), R2w, t2w and KF2's intrinsics K4_2.  sf2 / sigma2_2 are KF2's tables.
//   match12[i1] out = idx2 or -1 (vMatchedPairs = the non-negative entries in index order).
int match2_oracle_search_for_triangulation(
    const void* kps1, const uint8_t* desc1, int n1, const uint8_t* has_point1, int nn1, const int32_t* node1,
    const int32_t* start1, const int32_t* idx1, const void* kps2, const uint8_t* desc2, int n2,
    const uint8_t* has_point2, int nn2, const int32_t* node2, const int32_t* start2, const int32_t* idx2,
    const double* F12, const double* Cw, const double* R2w, const double* t2w, const float* K4_2, const float* sf2,
    const float* sigma2_2, int check_ori, int32_t* match12) {
  const KeyPoint* K1 = (const KeyPoint*)kps1;
  const KeyPoint* K2 = (const KeyPoint*)kps2;
  double C2[3];
  mat3_vec(R2w, Cw, t2w, C2);
  const float invz = 1.0f / (float)C2[2];
  const float ex = (float)(K4_2[0] * C2[0] * invz + K4_2[2]);
  const float ey = (float)(K4_2[1] * C2[1] * invz + K4_2[3]);
  const FeatVec A{nn1, node1, start1, idx1}, B{nn2, node2, start2, idx2};
  std::vector<int> rot_hist[kHistoLength];
  for (int i = 0; i < n1; i++) match12[i] = -1;
  int nmatches = 0, a = 0, b = 0;
  while (a < A.nn && b < B.nn) {
    if (A.node[a] == B.node[b]) {
      for (int k1 = A.start[a]; k1 < A.start[a + 1]; k1++) {
        const int i1 = A.idx[k1];
        if (has_point1[i1]) continue;
        const KeyPoint& kp1 = K1[i1];
        int best = kThLow, best_i2 = -1;
        for (int k2 = B.start[b]; k2 < B.start[b + 1]; k2++) {
          const int i2 = B.idx[k2];
          if (has_point2[i2]) continue;     // vbMatched2 is never set in this fork (:679-693)
          const int d = hamming256(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
          if (d > kThLow || d > best) continue;
          const KeyPoint& kp2 = K2[i2];
          const float distex = ex - kp2.x, distey = ey - kp2.y;
          if (distex * distex + distey * distey < 100 * sf2[kp2.octave]) continue;
          // CheckDistEpipolarLine :128-149 (float = float * double: evaluated in double, rounded on assignment)
          const float la = (float)(kp1.x * F12[0] + kp1.y * F12[3] + F12[6]);
          const float lb = (float)(kp1.x * F12[1] + kp1.y * F12[4] + F12[7]);
          const float lc = (float)(kp1.x * F12[2] + kp1.y * F12[5] + F12[8]);
          const float num = la * kp2.x + lb * kp2.y + lc;
          const float den = la * la + lb * lb;
          if (den == 0) continue;
          const float dsqr = num * num / den;
          if (dsqr < 3.84 * sigma2_2[kp2.octave]) { best_i2 = i2; best = d; }
        }
        if (best_i2 >= 0) {
          match12[i1] = best_i2;
          nmatches++;
          if (check_ori) rot_hist[rot_bin(kp1.angle, K2[best_i2].angle)].push_back(i1);
        }
      }
      a++; b++;
    } else if (A.node[a] < B.node[b]) {
      a = (int)(std::lower_bound(A.node, A.node + A.nn, B.node[b]) - A.node);
    } else {
      b = (int)(std::lower_bound(B.node, B.node + B.nn, A.node[a]) - B.node);
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rot_hist, kHistoLength, i1, i2, i3);
    for (int h = 0; h < kHistoLength; h++)
      if (h != i1 && h != i2 && h != i3)
        for (int idx : rot_hist[h]) { match12[idx] = -1; nmatches--; }
  }
  return nmatches;
}

// ---- SearchForInitialization :363-468 ----------------------------------------------------------------------------
//   F1: keypoints + descriptors; F2: full view (grid).  prev_matched [n1][2] in/out (vbPrevMatched).
//   matches12[i1] out = i2 or -1.
int match2_oracle_search_for_initialization(const void* kps1, const uint8_t* desc1, int n1, const void* kps2,
                                            const uint8_t* desc2, int n2, const int32_t* gs2, const int32_t* gi2,
                                            const float* view2, float* prev_matched, int window_size, float nn_ratio,
                                            int check_ori, int32_t* matches12) {
  const KeyPoint* K1 = (const KeyPoint*)kps1;
  const View F2 = make_view(kps2, desc2, n2, gs2, gi2, view2, nullptr, nullptr);
  int nmatches = 0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<int> rot_hist[kHistoLength];
  std::vector<int> matched_distance(n2, INT_MAX), matches21(n2, -1), cand;
  for (int i1 = 0; i1 < n1; i1++) {
    const int level1 = K1[i1].octave;
    if (level1 > 0) continue;
    features_in_area(F2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window_size, level1, level1, cand);
    if (cand.empty()) continue;
    int best = INT_MAX, best2 = INT_MAX, best_i2 = -1;
    for (int i2 : cand) {
      const int d = hamming256(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
      if (matched_distance[i2] <= d) continue;
      if (d < best) { best2 = best; best = d; best_i2 = i2; }
      else if (d < best2) best2 = d;
    }
    if (best <= kThLow) {
      if (best < (float)best2 * nn_ratio) {
        if (matches21[best_i2] >= 0) { matches12[matches21[best_i2]] = -1; nmatches--; }
        matches12[i1] = best_i2;
        matches21[best_i2] = i1;
        matched_distance[best_i2] = best;
        nmatches++;
        if (check_ori) rot_hist[rot_bin(K1[i1].angle, F2.kps[best_i2].angle)].push_back(i1);
      }
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rot_hist, kHistoLength, i1, i2, i3);
    for (int h = 0; h < kHistoLength; h++)
      if (h != i1 && h != i2 && h != i3)
        for (int idx1 : rot_hist[h])
          if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
  }
  for (int i1 = 0; i1 < n1; i1++)
    if (matches12[i1] >= 0) { prev_matched[2 * i1] = F2.kps[matches12[i1]].x; prev_matched[2 * i1 + 1] = F2.kps[matches12[i1]].y; }
  return nmatches;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:256-315) for a batch of map points ---------------------------
//   Point p owns the descriptors obs_start[p] .. obs_start[p+1]-1 (rows of `desc`), in the iteration order of its
//   std::map<KeyFrame*, size_t> with bad keyframes already dropped.  best[p] = index inside the point's list of the
//   descriptor with the least median distance to the others (first wins ties), -1 for an empty list (the reference
//   returns without touching descriptor_).  median = sorted row [ (int)(0.5 * (N - 1)) ], the row includes its own 0.
void map_oracle_distinctive_descriptors(int n_points, const int32_t* obs_start, const uint8_t* desc, int32_t* best) {
  std::vector<int> dists;
  for (int p = 0; p < n_points; p++) {
    const int o = obs_start[p], N = obs_start[p + 1] - o;
    best[p] = -1;
    if (N <= 0) continue;
    int best_median = INT_MAX, best_index = 0;
    for (int i = 0; i < N; i++) {
      dists.resize(N);
      for (int j = 0; j < N; j++) dists[j] = i == j ? 0 : hamming256(desc + 32 * (size_t)(o + i), desc + 32 * (size_t)(o + j));
      std::sort(dists.begin(), dists.end());
      const int median = dists[(size_t)(0.5 * (N - 1))];
      if (median < best_median) { best_median = median; best_index = i; }
    }
    best[p] = best_index;
  }
}

// ---- MapPoint::UpdateNormalAndDepth (MapPoint.cc:335-378) for a batch of map points -----------------------------------
//   obs_kf: keyframe index of every observation (CSR by point, map order); Ow [K][3] camera centres; per point the
//   reference keyframe and the octave of its keypoint there; sf = scale_factors_.  Points without observations keep
//   their outputs untouched.
void map_oracle_update_normal_and_depth(int n_points, const int32_t* obs_start, const int32_t* obs_kf, const double* Ow,
                                        const double* pos, const int32_t* ref_kf, const int32_t* ref_level,
                                        const float* sf, int n_levels, double* normal, float* min_distance,
                                        float* max_distance) {
  for (int p = 0; p < n_points; p++) {
    const int o = obs_start[p], N = obs_start[p + 1] - o;
    if (N <= 0) continue;
    const double* X = pos + 3 * p;
    double nrm[3] = {0.0, 0.0, 0.0};
    int n = 0;
    for (int k = 0; k < N; k++) {
      const double* C = Ow + 3 * (size_t)obs_kf[o + k];
      const double d[3] = {X[0] - C[0], X[1] - C[1], X[2] - C[2]};
      const double len = norm3(d);
      for (int a = 0; a < 3; a++) nrm[a] = nrm[a] + d[a] / len;
      n++;
    }
    const double* C = Ow + 3 * (size_t)ref_kf[p];
    const double PC[3] = {X[0] - C[0], X[1] - C[1], X[2] - C[2]};
    const float dist = (float)norm3(PC);
    const float level_scale_factor = sf[ref_level[p]];
    max_distance[p] = dist * level_scale_factor;
    min_distance[p] = max_distance[p] / sf[n_levels - 1];
    for (int a = 0; a < 3; a++) normal[3 * p + a] = nrm[a] / n;
  }
}

}  // extern "C"
