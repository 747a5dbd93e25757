// TEST INFRASTRUCTURE — C harness around the reference's OWN, unmodified data classes and matcher
// (/root/reference/src/{Frame,KeyFrame,MapPoint,Map,KeyFrameDatabase,ORBmatcher}.cc), compiled against the OpenCV / Eigen /
// glog stand-ins of this directory.  It builds real ORB_SLAM2::Frame / KeyFrame / MapPoint objects from the same POD arrays
// the oracle (oracle/matcher_oracle.cpp, matcher2_oracle.cpp) and the CUDA path take, calls the reference's methods and
// translates the pointer results back to indices.  Entry-point shapes follow the oracle's so tests run them side by side.
//
// `private` / `protected` are lifted for THIS translation unit only (to fill members that the reference fills from its
// tracking threads); the reference sources themselves are compiled untouched.
// every standard / stand-in header first, so that only the reference's own class definitions are read with the keywords lifted
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
#include <opencv/cv.h>
#include <algorithm>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include "ORBVocabulary.h"
#include "ORBextractor.h"
#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#undef private
#undef protected

#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "MatEigenConverter.h"

namespace refarena { void begin(); void end(); }

// src/MatEigenConverter.cc:87-95, the one converter Frame.cc / KeyFrame.cc call (see the shadow header)
std::vector<cv::Mat> MatEigenConverter::toDescriptorVector(const cv::Mat& Descriptors) {
  std::vector<cv::Mat> vDesc;
  vDesc.reserve(Descriptors.rows);
  for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
  return vDesc;
}

using namespace ORB_SLAM2;

namespace {

struct KP28 { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(KP28) == sizeof(cv::KeyPoint), "cv::KeyPoint layout");

// Frame's camera / grid parameters are static members (Frame.cc:127-147 computes them once per calibration)
void set_frame_statics(const float* bounds6, const float* K4) {
  Frame::min_x_ = bounds6[0]; Frame::max_x_ = bounds6[1]; Frame::min_y_ = bounds6[2]; Frame::max_y_ = bounds6[3];
  Frame::grid_element_width_inv_ = bounds6[4]; Frame::grid_element_height_inv_ = bounds6[5];
  if (K4) {
    Frame::fx_ = K4[0]; Frame::fy_ = K4[1]; Frame::cx_ = K4[2]; Frame::cy_ = K4[3];
    Frame::invfx_ = 1.0f / K4[0]; Frame::invfy_ = 1.0f / K4[1];
  }
  Frame::do_initial_computations_ = false;
}

void fill_scales(Frame& F, const float* scale_factors, int n_levels) {
  F.n_scale_levels_ = n_levels;
  F.scale_factors_.assign(scale_factors, scale_factors + n_levels);
  F.scale_factor_ = n_levels > 1 ? scale_factors[1] : 1.2f;
  F.log_scale_factor_ = std::log(F.scale_factor_);
  F.inv_scale_factors_.resize(n_levels); F.level_sigma2s_.resize(n_levels); F.inv_level_sigma2s_.resize(n_levels);
  for (int i = 0; i < n_levels; i++) {
    F.inv_scale_factors_[i] = 1.0f / scale_factors[i];
    F.level_sigma2s_[i] = scale_factors[i] * scale_factors[i];
    F.inv_level_sigma2s_[i] = 1.0f / F.level_sigma2s_[i];
  }
}

// A Frame as Tracking would hold it after Frame::Frame: keypoints, descriptors, the grid (the reference's own
// AssignFeaturesToGrid), no map points yet.
void fill_frame(Frame& F, const void* kps28, const uint8_t* desc, int n, const float* scale_factors, int n_levels) {
  F.N_ = n;
  F.undistort_keypoints_.resize(n);
  if (n) std::memcpy((void*)F.undistort_keypoints_.data(), kps28, (size_t)n * sizeof(KP28));
  F.keypoints_ = F.undistort_keypoints_;
  F.descriptors_.create(std::max(n, 1), 32, CV_8U);
  if (desc && n) std::memcpy(F.descriptors_.data, desc, (size_t)n * 32);
  F.mvuRight.assign(n, -1.0f);
  F.depthes_.assign(n, -1.0f);
  F.map_points_.assign(n, static_cast<MapPoint*>(nullptr));
  F.is_outliers_.assign(n, false);
  fill_scales(F, scale_factors, n_levels);
  F.AssignFeaturesToGrid();
}

void set_pose_rowmajor(Frame& F, const double* Tcw16) {
  Eigen::Matrix4d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = Tcw16[4 * r + c];
  F.SetPose(T);
}

// Owns the MapPoints of one call.  Points are created through the reference's Frame-based constructor (MapPoint.cc:71-110)
// from a one-keypoint seed frame, then given the state the caller describes.
struct PointPool {
  Map map;
  Frame seed;
  std::vector<std::unique_ptr<MapPoint> > pts;
  std::map<MapPoint*, int> index;
  PointPool(const float* scale_factors, int n_levels) {
    KP28 kp = {0, 0, 31, 0, 0, 0, -1};
    uint8_t d[32] = {0};
    const float b6[6] = {Frame::min_x_, Frame::max_x_, Frame::min_y_, Frame::max_y_, Frame::grid_element_width_inv_,
                         Frame::grid_element_height_inv_};
    (void)b6;
    fill_frame(seed, &kp, d, 1, scale_factors, n_levels);
    seed.SetPose(Eigen::Matrix4d::Identity());
  }
  MapPoint* add(const double* xw, const uint8_t* desc32, int n_obs) {
    Eigen::Vector3d pos(xw[0], xw[1], xw[2]);
    if (pos.norm() == 0.0) pos = Eigen::Vector3d(0, 0, 1e-30);   // the constructor normalises (pos - Ow); overwritten below
    const int idx = 0;
    MapPoint* p = new MapPoint(pos, &map, &seed, idx);
    p->world_pose_ = Eigen::Vector3d(xw[0], xw[1], xw[2]);
    if (desc32) { p->descriptor_.create(1, 32, CV_8U); std::memcpy(p->descriptor_.data, desc32, 32); }
    p->n_observations_ = n_obs;
    index[p] = (int)pts.size();
    pts.emplace_back(p);
    return p;
  }
  int index_of(MapPoint* p) const {
    if (!p) return -1;
    std::map<MapPoint*, int>::const_iterator it = index.find(p);
    return it == index.end() ? -2 : it->second;
  }
};

void frame_grid_csr(const Frame& F, int32_t* grid_start, int32_t* grid_idx) {
  int pos = 0;
  for (int i = 0; i < FRAME_GRID_COLS; i++)
    for (int j = 0; j < FRAME_GRID_ROWS; j++) {
      grid_start[i * FRAME_GRID_ROWS + j] = pos;
      for (size_t k = 0; k < F.grid_[i][j].size(); k++) grid_idx[pos++] = (int32_t)F.grid_[i][j][k];
    }
  grid_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = pos;
}

}  // namespace

extern "C" {

// Frame::Frame(imgGray, timestamp, extractor, voc, K, distCoef, bf, thDepth) — Frame.cc:91-156: ORB extraction,
// UndistortKeyPoints, ComputeImageBounds, grid pitch, AssignFeaturesToGrid, all by the reference's code.
// Outputs sized by the caller for `cap` keypoints; returns N_.
int ref_frame_construct(const uint8_t* img, int w, int h, int pitch, int nfeatures, float scale, int nlevels, int ini_th,
                        int min_th, const float* K4, const float* dist, int n_dist, int cap, void* kps28, void* un_kps28,
                        uint8_t* desc32, float* bounds6, int32_t* grid_start, int32_t* grid_idx) {
  ORBextractor extractor(nfeatures, scale, nlevels, ini_th, min_th);
  cv::Mat K(3, 3, CV_32F);
  for (int i = 0; i < 9; i++) K.at<float>(i / 3, i % 3) = 0.f;
  K.at<float>(0, 0) = K4[0]; K.at<float>(1, 1) = K4[1]; K.at<float>(0, 2) = K4[2]; K.at<float>(1, 2) = K4[3]; K.at<float>(2, 2) = 1.f;
  cv::Mat D(std::max(n_dist, 4), 1, CV_32F);
  for (int i = 0; i < std::max(n_dist, 4); i++) D.at<float>(i) = i < n_dist ? dist[i] : 0.f;
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)pitch);
  Frame::do_initial_computations_ = true;
  int n = 0;
  refarena::begin();     // the quadtree's address-order tie-break (ORBextractor.cc:684) needs ascending node addresses
  {
    Frame F(image, 0.0, &extractor, static_cast<ORBVocabulary*>(nullptr), K, D, 0.0f, 0.0f);
    n = F.N_;
    if (n <= cap) {
      if (n) {
        std::memcpy(kps28, F.keypoints_.data(), (size_t)n * sizeof(KP28));
        std::memcpy(un_kps28, F.undistort_keypoints_.data(), (size_t)n * sizeof(KP28));
        for (int i = 0; i < n; i++) std::memcpy(desc32 + 32 * (size_t)i, F.descriptors_.ptr(i), 32);
      }
      bounds6[0] = Frame::min_x_; bounds6[1] = Frame::max_x_; bounds6[2] = Frame::min_y_; bounds6[3] = Frame::max_y_;
      bounds6[4] = Frame::grid_element_width_inv_; bounds6[5] = Frame::grid_element_height_inv_;
      frame_grid_csr(F, grid_start, grid_idx);
    }
  }
  refarena::end();
  return n;
}

// Frame::AssignFeaturesToGrid (Frame.cc:158-173) + the CSR view of grid_
int ref_build_grid(const void* kps28, int n, const float* bounds6, const float* scale_factors, int n_levels, int32_t* grid_start,
                   int32_t* grid_idx) {
  set_frame_statics(bounds6, nullptr);
  Frame F;
  fill_frame(F, kps28, nullptr, n, scale_factors, n_levels);
  frame_grid_csr(F, grid_start, grid_idx);
  return grid_start[FRAME_GRID_COLS * FRAME_GRID_ROWS];
}

// Frame::GetFeaturesInArea (Frame.cc:243-307)
int ref_features_in_area(const void* kps28, int n, const float* bounds6, const float* scale_factors, int n_levels, float x,
                         float y, float r, int min_level, int max_level, int32_t* out) {
  set_frame_statics(bounds6, nullptr);
  Frame F;
  fill_frame(F, kps28, nullptr, n, scale_factors, n_levels);
  const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size(); i++) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th) — ORBmatcher.cc:1161-1271.
// Arguments as oracle/matcher_oracle.cpp: match_oracle_search_by_projection_frame.
int ref_search_by_projection_frame(const void* cur_kps28, const uint8_t* cur_desc, int cur_n, const float* bounds6,
                                   const float* K4, const float* scale_factors, int n_levels, const double* Tcw,
                                   const void* last_kps28, int last_n, const uint8_t* last_flags, const double* last_xw,
                                   const uint8_t* last_mp_desc, float th, int check_orientation, float nn_ratio,
                                   uint8_t* cur_claimed, int32_t* cur_match) {
  set_frame_statics(bounds6, K4);
  Frame cur, last;
  fill_frame(cur, cur_kps28, cur_desc, cur_n, scale_factors, n_levels);
  set_pose_rowmajor(cur, Tcw);
  fill_frame(last, last_kps28, nullptr, last_n, scale_factors, n_levels);
  last.SetPose(Eigen::Matrix4d::Identity());
  PointPool pool(scale_factors, n_levels);
  const double zero3[3] = {0, 0, 1};
  for (int i = 0; i < cur_n; i++)         // keypoints that already hold an observed map point (TrackReferenceKeyFrame leftovers)
    if (cur_claimed[i]) cur.map_points_[i] = pool.add(zero3, nullptr, 1);
  const int n_pre = (int)pool.pts.size();
  for (int i = 0; i < last_n; i++)
    if (last_flags[i] & 1) last.map_points_[i] = pool.add(last_xw + 3 * i, last_mp_desc + 32 * (size_t)i, (last_flags[i] >> 1) & 1);
  std::map<MapPoint*, int> last_index;
  for (int i = 0; i < last_n; i++) if (last.map_points_[i]) last_index[last.map_points_[i]] = i;
  ORBmatcher matcher(nn_ratio, check_orientation != 0);
  const int nm = matcher.SearchByProjection(cur, last, th);
  (void)n_pre;
  for (int i = 0; i < cur_n; i++) {
    MapPoint* p = cur.map_points_[i];
    std::map<MapPoint*, int>::const_iterator it = p ? last_index.find(p) : last_index.end();
    cur_match[i] = it == last_index.end() ? -1 : it->second;
    cur_claimed[i] = p && p->Observations() > 0;
  }
  return nm;
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) — ORBmatcher.cc:42-126.
// Arguments as match_oracle_search_by_projection_points; the per-point tracking fields are what Frame::isInFrustum left.
int ref_search_by_projection_points(const void* kps28, const uint8_t* desc, int n, const float* bounds6,
                                    const float* scale_factors, int n_levels, int n_points, const uint8_t* in_view,
                                    const int32_t* level, const float* view_cos, const float* proj_xy, const uint8_t* mp_desc,
                                    const uint8_t* has_obs, float th, float nn_ratio, uint8_t* claimed, int32_t* assign) {
  set_frame_statics(bounds6, nullptr);
  Frame F;
  fill_frame(F, kps28, desc, n, scale_factors, n_levels);
  PointPool pool(scale_factors, n_levels);
  const double zero3[3] = {0, 0, 1};
  for (int i = 0; i < n; i++)
    if (claimed[i]) F.map_points_[i] = pool.add(zero3, nullptr, 1);
  std::vector<MapPoint*> pts(n_points);
  std::map<MapPoint*, int> pidx;
  for (int p = 0; p < n_points; p++) {
    MapPoint* mp = pool.add(zero3, mp_desc + 32 * (size_t)p, has_obs[p] ? 1 : 0);
    mp->is_track_in_view_ = in_view[p] != 0;
    mp->track_scale_level_ = level[p];
    mp->track_view_cos_ = view_cos[p];
    mp->track_proj_x_ = proj_xy[2 * p]; mp->track_proj_y_ = proj_xy[2 * p + 1];
    mp->track_proj_x_r_ = -1.0f;
    pts[p] = mp; pidx[mp] = p;
  }
  ORBmatcher matcher(nn_ratio, true);
  const int nm = matcher.SearchByProjection(F, pts, th);
  for (int i = 0; i < n; i++) {
    MapPoint* p = F.map_points_[i];
    std::map<MapPoint*, int>::const_iterator it = p ? pidx.find(p) : pidx.end();
    assign[i] = it == pidx.end() ? -1 : it->second;
    claimed[i] = p && p->Observations() > 0;
  }
  return nm;
}

// Frame::isInFrustum (Frame.cc:191-241) with MapPoint::PredictScale (MapPoint.cc:405-420) for a batch of points.
// Arguments as match_oracle_is_in_frustum (pose15 = Rcw row-major, tcw, Ow).
void ref_is_in_frustum(const double* pose15, const float* K4, const float* bounds4, const float* scale_factors, int n_levels,
                       float cos_limit, int n_points, const double* xw, const double* normal, const float* min_dist,
                       const float* max_dist, uint8_t* in_view, float* proj_xy, int32_t* level, float* view_cos) {
  const float b6[6] = {bounds4[0], bounds4[1], bounds4[2], bounds4[3], 1.0f, 1.0f};
  set_frame_statics(b6, K4);
  Frame F;
  KP28 none = {0, 0, 31, 0, 0, 0, -1};
  fill_frame(F, &none, nullptr, 0, scale_factors, n_levels);
  Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T(r, c) = pose15[3 * r + c]; T(r, 3) = pose15[9 + r]; }
  F.SetPose(T);
  PointPool pool(scale_factors, n_levels);
  for (int p = 0; p < n_points; p++) {
    MapPoint* mp = pool.add(xw + 3 * p, nullptr, 1);
    mp->normal_vector_ = Eigen::Vector3d(normal[3 * p], normal[3 * p + 1], normal[3 * p + 2]);
    mp->min_distance_ = min_dist[p]; mp->max_distance_ = max_dist[p];
    const bool ok = F.isInFrustum(mp, cos_limit);
    in_view[p] = ok ? 1 : 0;
    if (ok) {
      proj_xy[2 * p] = mp->track_proj_x_; proj_xy[2 * p + 1] = mp->track_proj_y_;
      level[p] = mp->track_scale_level_; view_cos[p] = mp->track_view_cos_;
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// The keyframe searches (ORBmatcher.cc:128-1159, 1273-1384).  Views: kps / desc / n, bounds6 as the Frame holds them
// (KeyFrame::KeyFrame truncates them to int itself, KeyFrame.cc:80-83), K4, scale factors.

}  // extern "C"

namespace {

struct ViewArgs { const void* kps; const uint8_t* desc; int n; const float* bounds6; const float* K4; const float* sf; int nl; };

void view_frame(Frame& F, const ViewArgs& v, const double* Tcw16) {
  set_frame_statics(v.bounds6, v.K4);
  fill_frame(F, v.kps, v.desc, v.n, v.sf, v.nl);
  if (Tcw16) set_pose_rowmajor(F, Tcw16); else F.SetPose(Eigen::Matrix4d::Identity());
}
// pose12 = R (row-major 9) + t (3)
void pose12_to_T(const double* p, double* T16) {
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T16[4 * r + c] = p[3 * r + c]; T16[4 * r + 3] = p[9 + r]; }
  T16[12] = T16[13] = T16[14] = 0; T16[15] = 1;
}
void set_feature_vector(DBoW2::FeatureVector& fv, int nn, const int32_t* node, const int32_t* start, const int32_t* idx) {
  fv.clear();
  for (int k = 0; k < nn; k++)
    for (int e = start[k]; e < start[k + 1]; e++) fv.addFeature((DBoW2::NodeId)node[k], (unsigned)idx[e]);
}
Eigen::Matrix4d mat4_rowmajor(const double* T16) {
  Eigen::Matrix4d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = T16[4 * r + c];
  return T;
}
MapPoint* add_point_full(PointPool& pool, const double* xw, const double* normal, float min_d, float max_d, const uint8_t* desc) {
  MapPoint* p = pool.add(xw, desc, 0);
  if (normal) p->normal_vector_ = Eigen::Vector3d(normal[0], normal[1], normal[2]);
  p->min_distance_ = min_d; p->max_distance_ = max_d;
  return p;
}

}  // namespace

extern "C" {

// SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) — :1273-1384
int ref_search_by_projection_reloc(const void* kps28, const uint8_t* desc, int n, const float* bounds6, const float* K4,
                                   const float* sf, int nl, const double* Tcw, int n_kf, const uint8_t* kf_valid,
                                   const double* kf_xw, const float* kf_min_d, const float* kf_max_d, const uint8_t* kf_desc,
                                   const float* kf_angle, float th, int orb_dist, int check_ori, uint8_t* cur_has_point,
                                   int32_t* cur_match) {
  const ViewArgs v = {kps28, desc, n, bounds6, K4, sf, nl};
  Frame cur;
  view_frame(cur, v, Tcw);
  PointPool pool(sf, nl);
  const double z3[3] = {0, 0, 1};
  for (int i = 0; i < n; i++) if (cur_has_point[i]) cur.map_points_[i] = pool.add(z3, nullptr, 1);
  // the keyframe: only its keypoint angles and map points are read
  std::vector<KP28> kk(n_kf);
  for (int i = 0; i < n_kf; i++) { KP28 k = {0, 0, 31, kf_angle[i], 0, 0, -1}; kk[i] = k; }
  std::vector<uint8_t> kd((size_t)std::max(n_kf, 1) * 32, 0);
  Frame kf_frame;
  const ViewArgs kv = {kk.data(), kd.data(), n_kf, bounds6, K4, sf, nl};
  view_frame(kf_frame, kv, nullptr);
  KeyFrame kf(kf_frame, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  std::map<MapPoint*, int> kidx;
  for (int i = 0; i < n_kf; i++)
    if (kf_valid[i]) {
      MapPoint* p = add_point_full(pool, kf_xw + 3 * i, nullptr, kf_min_d[i], kf_max_d[i], kf_desc + 32 * (size_t)i);
      kf.AddMapPoint(p, i); kidx[p] = i;
    }
  std::set<MapPoint*> found;
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int nm = matcher.SearchByProjection(cur, &kf, found, th, orb_dist);
  for (int i = 0; i < n; i++) {
    MapPoint* p = cur.map_points_[i];
    std::map<MapPoint*, int>::const_iterator it = p ? kidx.find(p) : kidx.end();
    cur_match[i] = it == kidx.end() ? -1 : it->second;
    cur_has_point[i] = p != nullptr;
  }
  return nm;
}

// SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th) — :258-361
int ref_search_by_projection_sim3(const void* kps28, const uint8_t* desc, int n, const float* bounds6, const float* K4,
                                  const float* sf, int nl, const double* Scw, int n_points, const uint8_t* pt_skip,
                                  const double* xw, const double* normal, const float* min_d, const float* max_d,
                                  const uint8_t* pt_desc, int th, uint8_t* matched, int32_t* assign) {
  const ViewArgs v = {kps28, desc, n, bounds6, K4, sf, nl};
  Frame f;
  view_frame(f, v, nullptr);
  PointPool pool(sf, nl);
  KeyFrame kf(f, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  const double z3[3] = {0, 0, 1};
  std::vector<MapPoint*> vpMatched(n, static_cast<MapPoint*>(nullptr));
  for (int i = 0; i < n; i++) if (matched[i]) vpMatched[i] = pool.add(z3, nullptr, 1);
  std::vector<MapPoint*> pts(n_points);
  std::map<MapPoint*, int> pidx;
  for (int p = 0; p < n_points; p++) {
    pts[p] = add_point_full(pool, xw + 3 * p, normal + 3 * p, min_d[p], max_d[p], pt_desc + 32 * (size_t)p);
    if (pt_skip[p]) pts[p]->is_bad_ = true;
    pidx[pts[p]] = p;
  }
  ORBmatcher matcher(0.75f, true);
  const int nm = matcher.SearchByProjection(&kf, mat4_rowmajor(Scw), pts, vpMatched, th);
  for (int i = 0; i < n; i++) {
    std::map<MapPoint*, int>::const_iterator it = vpMatched[i] ? pidx.find(vpMatched[i]) : pidx.end();
    assign[i] = it == pidx.end() ? -1 : it->second;
    matched[i] = vpMatched[i] != nullptr;
  }
  return nm;
}

// Fuse(KeyFrame*, vpMapPoints, th) :724-842 (sim3 == 0, pose15 = Rcw, tcw, Ow) and Fuse(KeyFrame*, Scw, vpPoints, th,
// vpReplacePoint) :844-954 (sim3 == 1).  The keyframe starts without map points and every candidate is a distinct point, so
// the FIRST point that decides for a keypoint is added to the keyframe there (AddMapPoint) — that index is reported in
// added_idx[p]; later points deciding for the same keypoint go through Replace.  Returns nFused.
int ref_fuse(const void* kps28, const uint8_t* desc, int n, const float* bounds6, const float* K4, const float* sf, int nl,
             int sim3, const double* pose, int n_points, const uint8_t* pt_skip, const double* xw, const double* normal,
             const float* min_d, const float* max_d, const uint8_t* pt_desc, float th, int32_t* added_idx,
             int32_t* replace_idx) {
  const ViewArgs v = {kps28, desc, n, bounds6, K4, sf, nl};
  Frame f;
  double T16[16];
  if (!sim3) { pose12_to_T(pose, T16); view_frame(f, v, T16); } else view_frame(f, v, nullptr);
  PointPool pool(sf, nl);
  KeyFrame kf(f, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  std::vector<MapPoint*> pts(n_points);
  for (int p = 0; p < n_points; p++) {
    pts[p] = add_point_full(pool, xw + 3 * p, normal + 3 * p, min_d[p], max_d[p], pt_desc + 32 * (size_t)p);
    if (pt_skip[p]) pts[p]->is_bad_ = true;
    pool.map.AddMapPoint(pts[p]);
  }
  ORBmatcher matcher(0.6f, true);
  int nf;
  std::vector<MapPoint*> replace(n_points, static_cast<MapPoint*>(nullptr));
  if (sim3) nf = matcher.Fuse(&kf, mat4_rowmajor(pose), pts, th, replace);
  else nf = matcher.Fuse(&kf, pts, th);
  std::map<MapPoint*, int> pidx;
  for (int p = 0; p < n_points; p++) pidx[pts[p]] = p;
  for (int p = 0; p < n_points; p++) {
    added_idx[p] = (!pt_skip[p] && pts[p]->IsInKeyFrame(&kf)) ? pts[p]->GetIndexInKeyFrame(&kf) : -1;
    replace_idx[p] = replace[p] ? pidx[replace[p]] : -1;
  }
  return nf;
}

// SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) — :956-1159.  already2 is what the reference derives from
// vpMatches12 (the keyframe-2 index of every pre-matched point): already1_idx2[i1] = that index, -1 = none, -2 = not matched.
int ref_search_by_sim3(const void* kps1, const uint8_t* desc1, int n1, const void* kps2, const uint8_t* desc2, int n2,
                       const float* bounds6, const float* K4, const float* sf, int nl, const double* pose1, const double* pose2,
                       float s12, const double* R12, const double* t12, const uint8_t* valid1, const int32_t* already1_idx2,
                       const double* xw1, const float* min_d1, const float* max_d1, const uint8_t* mpdesc1, const uint8_t* valid2,
                       const double* xw2, const float* min_d2, const float* max_d2, const uint8_t* mpdesc2, float th,
                       int32_t* match12) {
  const ViewArgs v1 = {kps1, desc1, n1, bounds6, K4, sf, nl}, v2 = {kps2, desc2, n2, bounds6, K4, sf, nl};
  double T1[16], T2[16];
  pose12_to_T(pose1, T1); pose12_to_T(pose2, T2);
  Frame f1, f2;
  view_frame(f1, v1, T1); view_frame(f2, v2, T2);
  PointPool pool(sf, nl);
  KeyFrame kf1(f1, &pool.map, static_cast<KeyFrameDatabase*>(nullptr)), kf2(f2, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  std::map<MapPoint*, int> idx2;
  for (int i = 0; i < n1; i++)
    if (valid1[i]) kf1.AddMapPoint(add_point_full(pool, xw1 + 3 * i, nullptr, min_d1[i], max_d1[i], mpdesc1 + 32 * (size_t)i), i);
  for (int i = 0; i < n2; i++)
    if (valid2[i]) {
      MapPoint* p = add_point_full(pool, xw2 + 3 * i, nullptr, min_d2[i], max_d2[i], mpdesc2 + 32 * (size_t)i);
      kf2.AddMapPoint(p, i); idx2[p] = i;
    }
  const double z3[3] = {0, 0, 1};
  std::vector<MapPoint*> vpMatches12(n1, static_cast<MapPoint*>(nullptr));
  for (int i = 0; i < n1; i++)
    if (already1_idx2[i] > -2) {
      MapPoint* dmy = pool.add(z3, nullptr, 1);
      if (already1_idx2[i] >= 0) dmy->observations_[&kf2] = (size_t)already1_idx2[i];
      vpMatches12[i] = dmy;
    }
  Eigen::Matrix3d R;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R(r, c) = R12[3 * r + c];
  ORBmatcher matcher(0.75f, true);
  const int nf = matcher.SearchBySim3(&kf1, &kf2, vpMatches12, s12, R, Eigen::Vector3d(t12[0], t12[1], t12[2]), th);
  for (int i = 0; i < n1; i++) {
    std::map<MapPoint*, int>::const_iterator it = vpMatches12[i] ? idx2.find(vpMatches12[i]) : idx2.end();
    match12[i] = it == idx2.end() ? -1 : it->second;
  }
  return nf;
}

// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) :148-256 (mode 0: match[idx of F] = idx of KF) and
// SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) :462-580 (mode 1: match[idx1] = idx2).
int ref_search_by_bow(int mode, const void* kps1, const uint8_t* desc1, int n1, const uint8_t* valid1, int nn1,
                      const int32_t* node1, const int32_t* start1, const int32_t* feat1, const void* kps2, const uint8_t* desc2,
                      int n2, const uint8_t* valid2, int nn2, const int32_t* node2, const int32_t* start2, const int32_t* feat2,
                      const float* bounds6, const float* K4, const float* sf, int nl, float nn_ratio, int check_ori,
                      int32_t* match) {
  const ViewArgs v1 = {kps1, desc1, n1, bounds6, K4, sf, nl}, v2 = {kps2, desc2, n2, bounds6, K4, sf, nl};
  Frame f1, f2;
  view_frame(f1, v1, nullptr); view_frame(f2, v2, nullptr);
  set_feature_vector(f1.feature_vector_, nn1, node1, start1, feat1);
  set_feature_vector(f2.feature_vector_, nn2, node2, start2, feat2);
  PointPool pool(sf, nl);
  KeyFrame kf1(f1, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  const double z3[3] = {0, 0, 1};
  std::map<MapPoint*, int> i1;
  for (int i = 0; i < n1; i++)
    if (valid1[i]) { MapPoint* p = pool.add(z3, nullptr, 1); kf1.AddMapPoint(p, i); i1[p] = i; }
  ORBmatcher matcher(nn_ratio, check_ori != 0);
  if (mode == 0) {
    std::vector<MapPoint*> out;
    const int nm = matcher.SearchByBoW(&kf1, f2, out);
    for (int i = 0; i < n2; i++) {
      std::map<MapPoint*, int>::const_iterator it = out[i] ? i1.find(out[i]) : i1.end();
      match[i] = it == i1.end() ? -1 : it->second;
    }
    return nm;
  }
  KeyFrame kf2(f2, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  std::map<MapPoint*, int> i2;
  for (int i = 0; i < n2; i++)
    if (valid2[i]) { MapPoint* p = pool.add(z3, nullptr, 1); kf2.AddMapPoint(p, i); i2[p] = i; }
  std::vector<MapPoint*> out;
  const int nm = matcher.SearchByBoW(&kf1, &kf2, out);
  for (int i = 0; i < n1; i++) {
    std::map<MapPoint*, int>::const_iterator it = out[i] ? i2.find(out[i]) : i2.end();
    match[i] = it == i2.end() ? -1 : it->second;
  }
  return nm;
}

// SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo = false) — :582-722 with CheckDistEpipolarLine :128-146
int ref_search_for_triangulation(const void* kps1, const uint8_t* desc1, int n1, const uint8_t* has1, int nn1, const int32_t* node1,
                                 const int32_t* start1, const int32_t* feat1, const void* kps2, const uint8_t* desc2, int n2,
                                 const uint8_t* has2, int nn2, const int32_t* node2, const int32_t* start2, const int32_t* feat2,
                                 const double* F12, const double* pose1, const double* pose2, const float* bounds6,
                                 const float* K4, const float* sf, int nl, int check_ori, int32_t* match12) {
  const ViewArgs v1 = {kps1, desc1, n1, bounds6, K4, sf, nl}, v2 = {kps2, desc2, n2, bounds6, K4, sf, nl};
  double T1[16], T2[16];
  pose12_to_T(pose1, T1); pose12_to_T(pose2, T2);
  Frame f1, f2;
  view_frame(f1, v1, T1); view_frame(f2, v2, T2);
  set_feature_vector(f1.feature_vector_, nn1, node1, start1, feat1);
  set_feature_vector(f2.feature_vector_, nn2, node2, start2, feat2);
  PointPool pool(sf, nl);
  KeyFrame kf1(f1, &pool.map, static_cast<KeyFrameDatabase*>(nullptr)), kf2(f2, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  const double z3[3] = {0, 0, 1};
  for (int i = 0; i < n1; i++) if (has1[i]) kf1.AddMapPoint(pool.add(z3, nullptr, 1), i);
  for (int i = 0; i < n2; i++) if (has2[i]) kf2.AddMapPoint(pool.add(z3, nullptr, 1), i);
  Eigen::Matrix3d F;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F(r, c) = F12[3 * r + c];
  std::vector<std::pair<size_t, size_t> > pairs;
  ORBmatcher matcher(0.6f, check_ori != 0);
  const int nm = matcher.SearchForTriangulation(&kf1, &kf2, F, pairs, false);
  for (int i = 0; i < n1; i++) match12[i] = -1;
  for (size_t k = 0; k < pairs.size(); k++) match12[pairs[k].first] = (int32_t)pairs[k].second;
  return nm;
}

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — :364-460
int ref_search_for_initialization(const void* kps1, const uint8_t* desc1, int n1, const void* kps2, const uint8_t* desc2, int n2,
                                  const float* bounds6, const float* K4, const float* sf, int nl, float* prev_matched,
                                  int window_size, float nn_ratio, int check_ori, int32_t* match12) {
  const ViewArgs v1 = {kps1, desc1, n1, bounds6, K4, sf, nl}, v2 = {kps2, desc2, n2, bounds6, K4, sf, nl};
  Frame f1, f2;
  view_frame(f1, v1, nullptr); view_frame(f2, v2, nullptr);
  std::vector<cv::Point2f> prev(n1);
  for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nn_ratio, check_ori != 0);
  const int nm = matcher.SearchForInitialization(f1, f2, prev, m12, window_size);
  for (int i = 0; i < n1; i++) { match12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
  return nm;
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:256-315) and MapPoint::UpdateNormalAndDepth (:335-378) on real
// MapPoint / KeyFrame objects.  Observation lists are CSR by point (obs_start), obs_kf ascending and distinct inside a
// point; the keyframes are constructed in one block in index order, so the reference's std::map<KeyFrame*, size_t>
// iterates a point's observations in list order.  out_desc[p] = descriptor_ after the call (zeros for an empty list).
void ref_map_point_maintenance(int n_points, const int32_t* obs_start, const int32_t* obs_kf, const uint8_t* obs_desc, int n_kf,
                               const double* Ow, const double* pos, const int32_t* ref_kf, const int32_t* ref_level,
                               const float* sf, int nl, const float* bounds6, const float* K4, uint8_t* out_desc,
                               double* normal, float* min_d, float* max_d) {
  set_frame_statics(bounds6, K4);
  // rows of every keyframe: one keypoint per observation it hosts
  std::vector<std::vector<int> > rows(n_kf);           // observation ids per keyframe
  const int n_obs = obs_start[n_points];
  std::vector<int> row_of(n_obs);
  for (int o = 0; o < n_obs; o++) { row_of[o] = (int)rows[obs_kf[o]].size(); rows[obs_kf[o]].push_back(o); }
  std::vector<int> obs_point(n_obs);
  for (int p = 0; p < n_points; p++) for (int o = obs_start[p]; o < obs_start[p + 1]; o++) obs_point[o] = p;
  PointPool pool(sf, nl);
  // raw block: keyframe k lives at block + k * sizeof(KeyFrame)  (ascending addresses == ascending index)
  void* block = ::operator new(sizeof(KeyFrame) * (size_t)std::max(n_kf, 1) + 64);
  KeyFrame* kfs = (KeyFrame*)(((uintptr_t)block + 63) & ~(uintptr_t)63);
  for (int k = 0; k < n_kf; k++) {
    const int m = (int)rows[k].size();
    std::vector<KP28> kk(std::max(m, 1));
    std::vector<uint8_t> dd((size_t)std::max(m, 1) * 32, 0);
    for (int r = 0; r < m; r++) {
      const int o = rows[k][r], p = obs_point[o];
      KP28 kp = {0, 0, 31, 0, 0, (ref_kf[p] == k) ? ref_level[p] : 0, -1};
      kk[r] = kp;
      std::memcpy(&dd[32 * (size_t)r], obs_desc + 32 * (size_t)o, 32);
    }
    Frame f;
    fill_frame(f, kk.data(), dd.data(), m, sf, nl);
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    for (int a = 0; a < 3; a++) T(a, 3) = -Ow[3 * k + a];      // R = I: Ow = -R' t
    f.SetPose(T);
    new (&kfs[k]) KeyFrame(f, &pool.map, static_cast<KeyFrameDatabase*>(nullptr));
  }
  const double z3[3] = {0, 0, 1};
  for (int p = 0; p < n_points; p++) {
    MapPoint* mp = pool.add(pos + 3 * p, nullptr, 0);
    mp->normal_vector_ = Eigen::Vector3d(normal[3 * p], normal[3 * p + 1], normal[3 * p + 2]);
    mp->min_distance_ = min_d[p]; mp->max_distance_ = max_d[p];
    mp->descriptor_.create(1, 32, CV_8U);
    std::memset(mp->descriptor_.data, 0, 32);
    for (int o = obs_start[p]; o < obs_start[p + 1]; o++) mp->observations_[&kfs[obs_kf[o]]] = (size_t)row_of[o];
    mp->n_observations_ = obs_start[p + 1] - obs_start[p];
    mp->reference_keyframe_ = n_kf > 0 ? &kfs[ref_kf[p]] : nullptr;
    mp->ComputeDistinctiveDescriptors();
    mp->UpdateNormalAndDepth();
    std::memcpy(out_desc + 32 * (size_t)p, mp->descriptor_.data, 32);
    const Eigen::Vector3d nv = mp->GetNormal();
    normal[3 * p] = nv(0); normal[3 * p + 1] = nv(1); normal[3 * p + 2] = nv(2);
    min_d[p] = mp->min_distance_; max_d[p] = mp->max_distance_;
  }
  (void)z3;
  for (int k = 0; k < n_kf; k++) kfs[k].~KeyFrame();
  ::operator delete(block);
}

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1422-1437)
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
  std::memcpy(ma.data, a, 32); std::memcpy(mb.data, b, 32);
  return ORBmatcher::DescriptorDistance(ma, mb);
}

}  // extern "C"
