// Drop-in for the per-frame searches of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:36-102) over POD views.
// Implemented on the GPU: SearchByProjection(Frame&, vector<MapPoint*>&, th) and
// SearchByProjection(Frame& cur, const Frame& last, th), plus DescriptorDistance.  The other nine searches
// (BoW, Sim3, Fuse, initialisation, triangulation) are the "next" rows of SURVEY.md §8f and keep using the
// reference's CPU code.
#ifndef ORB_SLAM2_CMOS_ORBMATCHER_H
#define ORB_SLAM2_CMOS_ORBMATCHER_H

#include "views.h"

namespace ORB_SLAM2 {

class ORBmatcher {
 public:
  static const int TH_LOW = CMOS_TH_LOW;
  static const int TH_HIGH = CMOS_TH_HIGH;
  static const int HISTO_LENGTH = CMOS_HISTO_LENGTH;

  ORBmatcher(float nnratio = 0.6f, bool checkOri = true, int max_keypoints = 4096, int max_points = 8192, int device = 0)
      : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
    cmos_match_params p = {1, max_keypoints, max_points, device};
    cmos_throw_if(cmos_match_create(&p, &h_), "ORBmatcher");
  }
  ~ORBmatcher() { cmos_match_destroy(h_); }
  ORBmatcher(const ORBmatcher&) = delete;
  ORBmatcher& operator=(const ORBmatcher&) = delete;

  // Hamming distance of two 256-bit descriptors (ORBmatcher.cc:1422-1437)
  static int DescriptorDistance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
  }

  // Track the local map: returns the number of matches, writes F.map_points (ORBmatcher.cc:42-119)
  int SearchByProjection(FrameView& F, const MapPointsView& pts, const float th = 3) {
    bind(F);
    std::vector<int32_t> assign(F.N, -1);
    int32_t n = pts.n, nm = 0;
    cmos_throw_if(cmos_match_search_by_projection_points(h_, &n, pts.in_view, pts.level, pts.view_cos, pts.proj_xy,
                                                         pts.descriptors, pts.has_obs, pts.n, th, mfNNratio,
                                                         F.claimed.data(), assign.data(), &nm, 0, nullptr),
                  "ORBmatcher::SearchByProjection(F, points)");
    for (int i = 0; i < F.N; i++)
      if (assign[i] >= 0) F.map_points[i] = assign[i];
    return nm;
  }

  // Track from the previous frame (ORBmatcher.cc:1161-1271); F.map_points[i] = index of the last-frame keypoint
  int SearchByProjection(FrameView& CurrentFrame, const LastFrameView& LastFrame, const float th) {
    bind(CurrentFrame);
    std::vector<int32_t> match(CurrentFrame.N, -1);
    int32_t nl = LastFrame.N, nm = 0;
    cmos_throw_if(cmos_match_search_by_projection_frame(h_, CurrentFrame.Tcw, LastFrame.undistort_keypoints, &nl,
                                                        LastFrame.flags, LastFrame.world_pos, LastFrame.descriptors,
                                                        LastFrame.N, th, mbCheckOrientation ? 1 : 0,
                                                        CurrentFrame.claimed.data(), match.data(), &nm, 0, nullptr),
                  "ORBmatcher::SearchByProjection(cur, last)");
    for (int i = 0; i < CurrentFrame.N; i++)
      if (match[i] >= 0) CurrentFrame.map_points[i] = match[i];
    return nm;
  }

  cmos_match_t handle() { return h_; }

 protected:
  void bind(FrameView& F) {
    if ((int)F.map_points.size() != F.N) F.map_points.assign(F.N, -1);
    if ((int)F.claimed.size() != F.N) F.claimed.assign(F.N, 0);
    int32_t n = F.N;
    cmos_throw_if(cmos_match_set_frames(h_, &F.camera, F.undistort_keypoints, F.descriptors, &n, 1, F.N, 0, nullptr),
                  "ORBmatcher: Frame grid");
  }
  cmos_match_t h_ = nullptr;
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
#endif
