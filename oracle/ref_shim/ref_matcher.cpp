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

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1422-1437)
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
  std::memcpy(ma.data, a, 32); std::memcpy(mb.data, b, 32);
  return ORBmatcher::DescriptorDistance(ma, mb);
}

}  // extern "C"
