// POD views that stand in for the reference's pointer graph at the drop-in boundary.
// A maintainer of b51/ceres_mono_orb_slam2 fills these from Frame / KeyFrame / MapPoint (INTEGRATION.md shows
// the few lines per call site); with OpenCV / Eigen available the same adapters accept cv::Mat / cv::KeyPoint
// because the layouts are identical (cmos_keypoint == cv::KeyPoint, 28 bytes; poses are row-major doubles).
#ifndef ORB_SLAM2_CMOS_VIEWS_H
#define ORB_SLAM2_CMOS_VIEWS_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../cmos_b200.h"

namespace ORB_SLAM2 {

typedef cmos_keypoint KeyPoint;   // same layout as cv::KeyPoint

struct ImageView {                // cv::Mat of type CV_8UC1
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0;
  size_t step = 0;
  bool empty() const { return !data || rows == 0 || cols == 0; }
};

struct DescriptorMat {            // cv::Mat(N, 32, CV_8U)
  std::vector<uint8_t> data;
  int rows = 0;
  static const int cols = 32;
  void create(int n) { rows = n; data.assign((size_t)n * 32, 0); }
  uint8_t* ptr(int r) { return data.data() + (size_t)r * 32; }
  const uint8_t* ptr(int r) const { return data.data() + (size_t)r * 32; }
};

// What ORBmatcher reads from a Frame (Frame.h:158-214) and writes back (map_points_ as indices).
struct FrameView {
  cmos_camera camera;                       // bounds, grid pitch, intrinsics, scale factors
  const KeyPoint* undistort_keypoints = nullptr;
  const uint8_t* descriptors = nullptr;     // N x 32
  int N = 0;
  double Tcw[16];                           // row-major Tcw_
  std::vector<int32_t> map_points;          // per keypoint: map-point id or -1  (F.map_points_)
  std::vector<uint8_t> claimed;             // per keypoint: map_points_[i] && Observations() > 0
  cmos_feature_vector feature_vector = {0, nullptr, nullptr, nullptr};   // F.feature_vector_ (SearchByBoW)
};

// The last frame as SearchByProjection(CurrentFrame, LastFrame, th) reads it (ORBmatcher.cc:1176-1200).
struct LastFrameView {
  const KeyPoint* undistort_keypoints = nullptr;
  int N = 0;
  const uint8_t* flags = nullptr;           // bit0: map point present and not outlier; bit1: Observations() > 0
  const double* world_pos = nullptr;        // N x 3, MapPoint::GetWorldPos()
  const uint8_t* descriptors = nullptr;     // N x 32, MapPoint::GetDescriptor()
};

// Local map points as SearchByProjection(F, vpMapPoints, th) reads them after Frame::isInFrustum.
struct MapPointsView {
  int n = 0;
  const uint8_t* in_view = nullptr;         // is_track_in_view_ && !isBad()
  const int32_t* level = nullptr;           // track_scale_level_
  const float* view_cos = nullptr;          // track_view_cos_
  const float* proj_xy = nullptr;           // track_proj_x_, track_proj_y_
  const uint8_t* descriptors = nullptr;     // n x 32
  const uint8_t* has_obs = nullptr;         // Observations() > 0
};

// A KeyFrame (or a Frame used as one side of a two-view search) as the remaining searches read it: keypoints,
// descriptors, per-keypoint map point data (GetMapPointMatches()), pose, BoW feature vector, scale tables.
struct KeyFrameView {
  cmos_camera camera;                       // the Frame's bounds; the KeyFrame's int truncation is applied inside
  bool is_keyframe = true;
  const KeyPoint* undistort_keypoints = nullptr;
  const uint8_t* descriptors = nullptr;     // N x 32
  int N = 0;
  const uint8_t* mp_valid = nullptr;        // per keypoint: pMP && !pMP->isBad()
  const uint8_t* mp_present = nullptr;      // per keypoint: GetMapPoint(i) != NULL (SearchForTriangulation)
  const double* mp_world_pos = nullptr;     // N x 3
  const float* mp_min_distance = nullptr;   // raw min_distance_ / max_distance_ (0.8 / 1.2 applied inside)
  const float* mp_max_distance = nullptr;
  const uint8_t* mp_descriptors = nullptr;  // N x 32, MapPoint::GetDescriptor()
  cmos_feature_vector feature_vector = {0, nullptr, nullptr, nullptr};
  double Rcw[9], tcw[3], Ow[3];             // GetRotation(), GetTranslation(), GetCameraCenter()
  const float* level_sigma2 = nullptr;      // level_sigma2s_
  const float* inv_level_sigma2 = nullptr;  // inv_level_sigma2s_
};

// Candidate map points of Fuse / SearchByProjection(KeyFrame*, Scw, ...).
struct PointsView {
  int n = 0;
  const uint8_t* skip = nullptr;            // the reference's "continue" predicates that do not depend on the search
  const double* world_pos = nullptr;        // n x 3
  const double* normal = nullptr;           // n x 3, GetNormal()
  const float* min_distance = nullptr;
  const float* max_distance = nullptr;
  const uint8_t* descriptors = nullptr;     // n x 32
};

// Fuse returns decisions: point p should be fused with keypoint best_idx[p] (or -1).  The caller walks them in
// order and applies Replace / AddObservation exactly as ORBmatcher.cc:822-836 / 938-947.
struct FuseDecisions {
  std::vector<int32_t> best_idx, best_dist;
};

// PoseOptimization(Frame*): the matched keypoints of one frame.
struct FramePoseView {
  double pose7[7];                          // Tcw_ as [t, q_xyzw] (MatEigenConverter::Matrix4dToMatrix_7_1)
  int n = 0;                                // keypoints with a map point
  const double* world_pos = nullptr;        // n x 3
  const float* uv = nullptr;                // n x 2, undistort_keypoints_[i].pt
  const float* inv_sigma2 = nullptr;        // n, inv_level_sigma2s_[octave]
  float K4[4];                              // fx, fy, cx, cy
  std::vector<uint8_t> is_outlier;          // out: is_outliers_
};

// LocalBundleAdjustment / BundleAdjustment: keyframes, map points, observations.
struct GraphView {
  int n_keyframes = 0, n_points = 0, n_obs = 0;
  double* keyframe_pose7 = nullptr;         // in/out, n_keyframes x 7
  const uint8_t* keyframe_flags = nullptr;  // bit0 constant (fixed keyframe or id 0), bit1 not a local keyframe
  double* point_pos = nullptr;              // in/out, n_points x 3
  const int32_t* obs_keyframe = nullptr;
  const int32_t* obs_point = nullptr;
  const float* obs_uv = nullptr;            // n_obs x 2
  const float* obs_inv_sigma2 = nullptr;
  float K4[4];
  std::vector<uint8_t> erase;               // out (LocalBA): observations to erase from the map
};

// The map as LocalBundleAdjustment(keyframe, stop_flag, map) and GlobalBundleAdjustemnt(map, ...) walk it
// (CeresOptimizer.cc:53-57, 349-406): keyframes and map points by index with the accessors those functions call.
struct MapView {
  int n_keyframes = 0, n_points = 0;
  std::vector<unsigned long> keyframe_id;                  // id_
  std::vector<uint8_t> keyframe_bad;                       // isBad()
  const double* keyframe_pose7 = nullptr;                  // n_keyframes x 7: Matrix4dToMatrix_7_1(GetPose())
  std::vector<std::vector<int> > covisible;                // GetVectorCovisibleKeyFrames(), indices, in its order
  std::vector<std::vector<int> > keyframe_points;          // GetMapPointMatches(): per keypoint a point index or -1
  std::vector<const KeyPoint*> undistort_keypoints;        // undistort_keypoints_ of every keyframe
  std::vector<const float*> inv_level_sigma2;              // inv_level_sigma2s_ of every keyframe
  std::vector<uint8_t> point_bad;                          // isBad()
  const double* point_pos = nullptr;                       // n_points x 3: GetWorldPos()
  std::vector<std::vector<std::pair<int, int> > > observations;   // GetObservations(): (keyframe index, keypoint index)
  float K4[4];                                             // fx_, fy_, cx_, cy_ (one camera)
};

// A bundle-adjustment problem collected from a MapView: owns the arrays a GraphView points at, and remembers which map
// keyframe / point every graph row is, so the result can be written back (SetPose / SetWorldPos / erase lists).
struct CollectedGraph {
  std::vector<int> keyframe_index, point_index;            // graph row -> map index
  std::vector<double> pose7, point_pos;
  std::vector<uint8_t> flags;                              // bit0 constant, bit1 not a local keyframe
  std::vector<int32_t> obs_keyframe, obs_point;            // graph rows
  std::vector<int> obs_keypoint;                           // keypoint index of the observation in its keyframe
  std::vector<float> obs_uv, obs_inv_sigma2;
  GraphView view(const float K4[4]) {
    GraphView g;
    g.n_keyframes = (int)keyframe_index.size(); g.n_points = (int)point_index.size(); g.n_obs = (int)obs_keyframe.size();
    g.keyframe_pose7 = pose7.data(); g.keyframe_flags = flags.data(); g.point_pos = point_pos.data();
    g.obs_keyframe = obs_keyframe.data(); g.obs_point = obs_point.data(); g.obs_uv = obs_uv.data();
    g.obs_inv_sigma2 = obs_inv_sigma2.data();
    for (int k = 0; k < 4; k++) g.K4[k] = K4[k];
    return g;
  }
};

// Sophus::Sim3d as scale(), rotationMatrix() (row-major), translation() — 13 doubles, the layout the C ABI reads.
struct Sim3POD {
  double s = 1.0;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double t[3] = {0, 0, 0};
};

// OptimizeSim3(keyframe_1, keyframe_2, matches12, S12, th2, bFixScale): the correspondences that pass the tests of
// CeresOptimizer.cc:646-654, in matches12 order.
struct Sim3MatchesView {
  int n = 0;
  float K1[4], K2[4];                       // fx, fy, cx, cy
  const float* obs1 = nullptr;              // n x 2, keyframe_1->undistort_keypoints_[i].pt
  const float* inv_sigma1 = nullptr;        // n, keyframe_1->inv_level_sigma2s_[octave]
  const double* P3D2c = nullptr;            // n x 3, R2cw * map_point_2 + t2cw
  const float* obs2 = nullptr;
  const float* inv_sigma2 = nullptr;
  const double* P3D1c = nullptr;
  std::vector<uint8_t> is_bad;              // out: is_outlier_12 || is_outlier_21 per correspondence
};

// OptimizeEssentialGraph(map, loop_keyframe, current_keyframe, non_corrected, corrected, loop_connections, fixed_scale):
// the map's non-bad keyframes (map->GetAllKeyFrames()) by index, with the accessors the function calls.
struct EssentialGraphView {
  int n_keyframes = 0;
  std::vector<unsigned long> id;                    // id_
  std::vector<Sim3POD> pose;                        // (1, GetRotation(), GetTranslation())
  std::vector<int> parent;                          // GetParent() or -1
  std::vector<std::vector<int> > children;          // hasChild
  std::vector<std::vector<int> > loop_edges;        // GetLoopEdges()
  std::vector<std::vector<int> > covisibles;        // GetCovisiblesByWeight(100), in its (weight-sorted) order
  std::vector<std::vector<int> > loop_connections;  // loop_connections[keyframe], empty when the keyframe has no entry
  std::vector<std::vector<int> > loop_connection_weight;   // GetWeight(keyframe, connected) of every entry above
  std::vector<uint8_t> has_corrected, has_non_corrected;
  std::vector<Sim3POD> corrected, non_corrected;    // keyframes_corrected_sim3 / keyframes_non_corrected_sim3 entries
  int loop_keyframe = -1, current_keyframe = -1;    // indices
  int n_points = 0;                                 // map->GetAllMapPoints() that are not bad
  const double* point_pos = nullptr;                // n_points x 3
  const int32_t* point_ref = nullptr;               // index of corrected_reference_ / GetReferenceKeyFrame() (:939-946)
  std::vector<double> Tiw;                          // out: n_keyframes x 16 for SetPose
  std::vector<double> corrected_pos;                // out: n_points x 3 for SetWorldPos
};

inline void cmos_throw_if(int status, const char* what) {
  if (status != CMOS_OK) throw std::runtime_error(std::string(what) + ": " + cmos_last_error());
}

}  // namespace ORB_SLAM2
#endif
